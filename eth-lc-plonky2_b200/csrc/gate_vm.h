// Gate-expression bytecode: how a circuit's custom gates reach the quotient kernel (row a6) and the verifier (row f4).
//
// plonky2 evaluates a gate through Gate::eval_unfiltered_base_batch (prover, base field) and Gate::eval_unfiltered
// (verifier, F_p^2)  [DEP plonky2:gates/gate.rs, plonk/vanishing_poly.rs::evaluate_gate_constraints(_base_batch)];
// the eth-lc circuit uses about twenty gate types from plonky2 and plonky2_crypto (/root/reference/eth-lc-plonky2/src/
// targets.rs:468-470 recursion gates, merkle_tree_gadget.rs:37 SHA-256 u32 gates, utils.rs:102-103 BaseSumGate).  Porting
// each gate by hand would tie the engine to one gate set, so a gate is DATA here: a straight-line program over
//     wires[i]  (local wire i)      consts[i]  (local constant i, after the selector prefix)
//     imm[i]    (64-bit immediates of the circuit; imm[0..4) = public_inputs_hash)        r0 .. r63 (registers)
// with the operations add / sub / mul / mad (a*b + c) / msub (a*b - c) / mov / emit.  `emit` yields the next constraint of
// the gate, in plonky2's order.  The Rust side obtains a program by driving the gate's eval_unfiltered_circuit (which
// plonky2 has for recursion) into a recorder; gate_lib.h builds the same programs for the gates restated
// in SURVEY.md A.8 and DESIGN.md.  One interpreter serves the device (u64, base field, quotient kernel) and the host
// (F_p^2, verifier): it is a template over the field.
//
// Instruction word (u64):  op[0,8)  dst[8,20)  a[20,34)  b[34,48)  c[48,62);  operand = space << 12 | index.
#pragma once
#include "gl64.cuh"
#include <map>
#include <string>
#include <vector>

#define GVM_NREG 64
enum GvmOp : u32 { GVM_END = 0, GVM_ADD = 1, GVM_SUB = 2, GVM_MUL = 3, GVM_MAD = 4, GVM_MSUB = 5, GVM_MOV = 6, GVM_EMIT = 7, GVM_NUM_OPS = 8 };
enum GvmSpace : u32 { GVM_REG = 0, GVM_WIRE = 1, GVM_CONST = 2, GVM_IMM = 3 };
#define GVM_NUM_PI 4   // imm[0..4) are the public-inputs hash (filled per proof)

GL_HD u64 gvm_pack(u32 op, u32 dst, u32 a, u32 b, u32 c) {
    return (u64)op | ((u64)dst << 8) | ((u64)a << 20) | ((u64)b << 34) | ((u64)c << 48);
}
GL_HD u32 gvm_operand(u32 space, u32 index) { return (space << 12) | index; }

// The interpreter.  F: field (T, add, sub, mul, mad, msub).  X: operand context with wire(i), constant(i), imm(i) -> F::T.
// regs: GVM_NREG values of scratch owned by the caller.  emit(v) receives the constraints in order.
template <class F, class X, class E>
GL_HD void gvm_run(const u64 *prog, u32 len, const X &cx, typename F::T *regs, E &emit) {
    typedef typename F::T T;
    for (u32 pc = 0; pc < len; pc++) {
        const u64 w = prog[pc];
        const u32 op = (u32)(w & 0xff), dst = (u32)(w >> 8) & 0xfff;
        const u32 sa = (u32)(w >> 20) & 0x3fff, sb = (u32)(w >> 34) & 0x3fff, sc = (u32)(w >> 48) & 0x3fff;
        if (op == GVM_END) break;
#define GVM_FETCH(s) (((s) >> 12) == GVM_REG ? regs[(s) & 0xfff] : ((s) >> 12) == GVM_WIRE ? cx.wire((s) & 0xfff) : ((s) >> 12) == GVM_CONST ? cx.constant((s) & 0xfff) : cx.imm((s) & 0xfff))
        const T a = GVM_FETCH(sa);
        if (op == GVM_EMIT) { emit(a); continue; }
        if (op == GVM_MOV) { regs[dst] = a; continue; }
        const T b = GVM_FETCH(sb);
        T r;
        if (op == GVM_ADD) r = F::add(a, b);
        else if (op == GVM_SUB) r = F::sub(a, b);
        else if (op == GVM_MUL) r = F::mul(a, b);
        else {
            const T c = GVM_FETCH(sc);
            r = (op == GVM_MAD) ? F::mad(a, b, c) : F::msub(a, b, c);
        }
#undef GVM_FETCH
        regs[dst] = r;
    }
}

struct GvmBaseField {
    typedef u64 T;
    static GL_HD T add(T a, T b) { return gl_add(a, b); }
    static GL_HD T sub(T a, T b) { return gl_sub(a, b); }
    static GL_HD T mul(T a, T b) { return gl_mul(a, b); }
    static GL_HD T mad(T a, T b, T c) { return gl_mul_add(a, b, c); }
    static GL_HD T msub(T a, T b, T c) { return gl_sub(gl_mul(a, b), c); }
};
struct GvmExtField {
    typedef gl2 T;
    static GL_HD T add(T a, T b) { return gl2_add(a, b); }
    static GL_HD T sub(T a, T b) { return gl2_sub(a, b); }
    static GL_HD T mul(T a, T b) { return gl2_mul(a, b); }
    static GL_HD T mad(T a, T b, T c) { return gl2_add(gl2_mul(a, b), c); }
    static GL_HD T msub(T a, T b, T c) { return gl2_sub(gl2_mul(a, b), c); }
};

// ---- host: validation of a program that arrived over the ABI (indices in range, registers written before read) ----
static inline bool gvm_validate(const u64 *prog, u32 len, u32 num_wires, u32 num_consts, u32 num_imm, u32 *num_constraints, std::string *why) {
    bool written[GVM_NREG] = {false};
    u32 emits = 0;
    auto bad = [&](const char *m, u32 pc) { if (why) *why = std::string(m) + " at instruction " + std::to_string(pc); return false; };
    for (u32 pc = 0; pc < len; pc++) {
        const u64 w = prog[pc];
        const u32 op = (u32)(w & 0xff), dst = (u32)(w >> 8) & 0xfff;
        const u32 s[3] = {(u32)(w >> 20) & 0x3fff, (u32)(w >> 34) & 0x3fff, (u32)(w >> 48) & 0x3fff};
        if (w >> 62) return bad("reserved bits set", pc);
        if (op == GVM_END) break;
        if (op >= GVM_NUM_OPS) return bad("unknown opcode", pc);
        const int nsrc = (op == GVM_EMIT || op == GVM_MOV) ? 1 : (op == GVM_MAD || op == GVM_MSUB) ? 3 : 2;
        for (int k = 0; k < nsrc; k++) {
            const u32 sp = s[k] >> 12, ix = s[k] & 0xfff;
            if (sp == GVM_REG) { if (ix >= GVM_NREG || !written[ix]) return bad("register read before write / out of range", pc); }
            else if (sp == GVM_WIRE) { if (ix >= num_wires) return bad("wire index out of range", pc); }
            else if (sp == GVM_CONST) { if (ix >= num_consts) return bad("constant index out of range", pc); }
            else if (ix >= num_imm) return bad("immediate index out of range", pc);
        }
        if (op == GVM_EMIT) { emits++; continue; }
        if (dst >= GVM_NREG) return bad("destination register out of range", pc);
        written[dst] = true;
    }
    if (num_constraints) *num_constraints = emits;
    return true;
}

// ---- host: program builder.  Values are SSA; finish() assigns the GVM_NREG physical registers by liveness. ----
struct GvmVal {
    u32 space, index;   // space 0: SSA value `index`; otherwise a direct operand
};
struct GvmImmPool {
    std::vector<u64> values;              // values[0..4) reserved for public_inputs_hash
    std::map<u64, u32> index;
    GvmImmPool() : values(GVM_NUM_PI, 0) {}
    u32 get(u64 v) {
        v = gl_canon(v);
        auto it = index.find(v);
        if (it != index.end()) return it->second;
        u32 i = (u32)values.size();
        values.push_back(v);
        index[v] = i;
        return i;
    }
};
struct GvmBuilder {
    typedef GvmVal V;                          // gate_lib.h is written against a builder concept: V + the methods below
    static constexpr bool kDirect = false;     // (QuotDirect in plonk.cuh is the other model: V = u64, evaluates in place)
    struct Ins { u32 op; u32 dst; GvmVal s[3]; int nsrc; };
    std::vector<Ins> ins;
    GvmImmPool &pool;
    u32 next_ssa = 0, num_emits = 0;
    bool ok = true;
    explicit GvmBuilder(GvmImmPool &p) : pool(p) {}

    GvmVal wire(u32 i) const { return GvmVal{GVM_WIRE, i}; }
    GvmVal constant(u32 i) const { return GvmVal{GVM_CONST, i}; }
    GvmVal pi(u32 i) const { return GvmVal{GVM_IMM, i}; }
    GvmVal imm(u64 v) { return GvmVal{GVM_IMM, pool.get(v)}; }
    GvmVal op3(u32 op, GvmVal a, GvmVal b, GvmVal c, int nsrc) {
        Ins in; in.op = op; in.dst = next_ssa++; in.s[0] = a; in.s[1] = b; in.s[2] = c; in.nsrc = nsrc;
        ins.push_back(in);
        return GvmVal{GVM_REG, in.dst};
    }
    GvmVal add(GvmVal a, GvmVal b) { return op3(GVM_ADD, a, b, a, 2); }
    GvmVal sub(GvmVal a, GvmVal b) { return op3(GVM_SUB, a, b, a, 2); }
    GvmVal mul(GvmVal a, GvmVal b) { return op3(GVM_MUL, a, b, a, 2); }
    GvmVal mad(GvmVal a, GvmVal b, GvmVal c) { return op3(GVM_MAD, a, b, c, 3); }
    GvmVal msub(GvmVal a, GvmVal b, GvmVal c) { return op3(GVM_MSUB, a, b, c, 3); }
    void emit(GvmVal a) {
        Ins in; in.op = GVM_EMIT; in.dst = 0; in.s[0] = a; in.s[1] = a; in.s[2] = a; in.nsrc = 1;
        ins.push_back(in);
        num_emits++;
    }
    // x^k for small k by repeated squaring
    GvmVal pow(GvmVal x, u32 k) {
        GvmVal r = x; bool have = false; GvmVal base = x;
        while (k) {
            if (k & 1) { r = have ? mul(r, base) : base; have = true; }
            k >>= 1;
            if (k) base = mul(base, base);
        }
        return r;
    }
    // sum_i v[i] * base^i  (reduce_with_powers), Horner from the top
    GvmVal reduce_with_powers(const GvmVal *v, u32 n, u64 base) {
        GvmVal b = imm(base), acc = v[n - 1];
        for (u32 i = n - 1; i-- > 0;) acc = mad(acc, b, v[i]);
        return acc;
    }
    // prod_{k < count} (x - k)
    GvmVal range_product(GvmVal x, u32 count) {
        if (count == 4) {   // x (x - 1)(x - 2)(x - 3) = y (y + 2), y = x (x - 3): two products instead of three (the u32 gates' limbs)
            GvmVal y = mul(x, sub(x, imm(3)));
            return mul(y, add(y, imm(2)));
        }
        GvmVal acc = x;   // (x - 0)
        for (u32 k = 1; k < count; k++) acc = mul(acc, sub(x, imm(k)));
        return acc;
    }

    // Physical registers: a value's register is released at its last use, BEFORE the destination of that instruction is
    // chosen (the interpreter reads all operands before it writes), so chains run in place.
    std::vector<u64> finish() {
        std::vector<int> last(next_ssa, -1);
        for (size_t k = 0; k < ins.size(); k++)
            for (int j = 0; j < ins[k].nsrc; j++)
                if (ins[k].s[j].space == GVM_REG) last[ins[k].s[j].index] = (int)k;
        std::vector<int> phys(next_ssa, -1);
        std::vector<bool> busy(GVM_NREG, false);
        std::vector<u64> out;
        for (size_t k = 0; k < ins.size(); k++) {
            const Ins &in = ins[k];
            u32 s[3] = {0, 0, 0};
            for (int j = 0; j < 3; j++) {
                const GvmVal &v = in.s[j < in.nsrc ? j : 0];
                s[j] = v.space == GVM_REG ? gvm_operand(GVM_REG, (u32)phys[v.index]) : gvm_operand(v.space, v.index);
            }
            for (int j = 0; j < in.nsrc; j++)
                if (in.s[j].space == GVM_REG && last[in.s[j].index] == (int)k) busy[phys[in.s[j].index]] = false;
            u32 dst = 0;
            if (in.op != GVM_EMIT) {
                if (last[in.dst] < 0) continue;    // dead value
                int r = -1;
                for (int q = 0; q < GVM_NREG; q++) if (!busy[q]) { r = q; break; }
                if (r < 0) { ok = false; return out; }   // more than GVM_NREG live values
                busy[r] = true; phys[in.dst] = r; dst = (u32)r;
            }
            out.push_back(gvm_pack(in.op, dst, s[0], in.nsrc > 1 ? s[1] : 0, in.nsrc > 2 ? s[2] : 0));
        }
        out.push_back(gvm_pack(GVM_END, 0, 0, 0, 0));
        return out;
    }
};

// F_p^2 = F_p[X]/(X^2 - 7) on top of a builder (the recursion gates work on extension elements spread over D = 2 wires)
template <class BT> struct GvmExtT { typename BT::V a, b; };
template <class BT> struct GvmExtOpsT {
    typedef typename BT::V V;
    typedef GvmExtT<BT> E;
    BT &B;
    GL_HD explicit GvmExtOpsT(BT &b) : B(b) {}
    GL_HD E wires(u32 first) { return E{B.wire(first), B.wire(first + 1)}; }
    GL_HD E add(E x, E y) { return E{B.add(x.a, y.a), B.add(x.b, y.b)}; }
    GL_HD E sub(E x, E y) { return E{B.sub(x.a, y.a), B.sub(x.b, y.b)}; }
    GL_HD E mul(E x, E y) {
        V bb7 = B.mul(B.mul(x.b, y.b), B.imm(7));
        return E{B.mad(x.a, y.a, bb7), B.mad(x.a, y.b, B.mul(x.b, y.a))};
    }
    GL_HD E scale(E x, V s) { return E{B.mul(x.a, s), B.mul(x.b, s)}; }
    GL_HD void emit(E x) { B.emit(x.a); B.emit(x.b); }
};
