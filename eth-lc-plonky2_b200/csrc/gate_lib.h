// Gate library: bytecode programs (gate_vm.h) of the plonky2 / plonky2_crypto gates whose constraint formulas are restated
// in SURVEY.md A.8 and DESIGN.md 6.  Host code.  [DEP plonky2:gates/{constant,public_input,arithmetic_base,poseidon,base_sum,
// arithmetic_extension,multiplication_extension,reducing,reducing_extension,random_access,exponentiation,poseidon_mds,
// coset_interpolation}.rs]
// [DEP plonky2_crypto (plonky2_u32):gates/{arithmetic_u32,add_many_u32,subtraction_u32,range_check_u32,comparison,
//      interleave_u32,uninterleave_to_u32,uninterleave_to_b32}.rs]
// Tier C of SURVEY.md Appendix A: wire layouts and constraint ORDER are recalled, not checked against the absent source;
// with a Rust toolchain the programs are RECORDED from the gates' own eval_unfiltered_circuit instead and
// these builders become test fixtures.  Constraint t of a gate receives alpha^t, so order is part of the contract.
#pragma once
#include "gate_vm.h"
#ifdef __CUDACC__
#pragma nv_diag_suppress 20011, 20014   // the builder-model templates are __host__ __device__; GvmBuilder is their host-only model
#endif
#include "poseidon_consts.h"

enum PlkGateKind : u32 {
    PLK_NOOP = 0, PLK_CONSTANT = 1, PLK_PUBLIC_INPUT = 2, PLK_ARITHMETIC = 3, PLK_POSEIDON = 4, PLK_BASE_SUM = 5,
    PLK_ARITHMETIC_EXT = 6, PLK_MUL_EXT = 7, PLK_REDUCING = 8, PLK_REDUCING_EXT = 9, PLK_RANDOM_ACCESS = 10,
    PLK_EXPONENTIATION = 11, PLK_POSEIDON_MDS = 12, PLK_U32_ARITHMETIC = 13, PLK_U32_ADD_MANY = 14, PLK_U32_SUBTRACTION = 15,
    PLK_U32_RANGE_CHECK = 16, PLK_COMPARISON = 17, PLK_COSET_INTERPOLATION = 18,
    PLK_U32_INTERLEAVE = 19, PLK_UNINTERLEAVE_TO_U32 = 20, PLK_UNINTERLEAVE_TO_B32 = 21, PLK_NUM_KINDS = 22,
    PLK_CUSTOM = 255   // no library builder: the program came over the ABI (recorded on the Rust side)
};

// highest constraint degree of the gate (drives plonky2's selector grouping; used by the synthetic circuit generator)
static inline u32 plk_gate_degree(u32 kind, const u32 p[4]) {
    switch (kind) {
        case PLK_NOOP: return 0;
        case PLK_CONSTANT: case PLK_PUBLIC_INPUT: case PLK_POSEIDON_MDS: return 1;
        case PLK_BASE_SUM: return p[0];
        case PLK_REDUCING: case PLK_REDUCING_EXT: return 2;
        case PLK_ARITHMETIC: case PLK_ARITHMETIC_EXT: case PLK_MUL_EXT: return 3;
        case PLK_U32_ARITHMETIC: case PLK_U32_ADD_MANY: case PLK_U32_SUBTRACTION: case PLK_U32_RANGE_CHECK: case PLK_EXPONENTIATION: return 4;
        case PLK_COMPARISON: return 1u << ((p[0] + p[1] - 1) / p[1]);
        case PLK_COSET_INTERPOLATION: return p[1];
        case PLK_U32_INTERLEAVE: case PLK_UNINTERLEAVE_TO_U32: case PLK_UNINTERLEAVE_TO_B32: return 2;
        case PLK_RANDOM_ACCESS: return p[0] + 1;
        case PLK_POSEIDON: return 7;
    }
    return 0;
}

// ---- wire layouts (shared with the synthetic witness filler) ----
struct U32ArithLayout {   // U32ArithmeticGate: 6 routed wires per op, then 32 two-bit limbs per op
    u32 num_ops;
    GL_HD u32 m0(u32 i) const { return 6 * i; }
    GL_HD u32 m1(u32 i) const { return 6 * i + 1; }
    GL_HD u32 addend(u32 i) const { return 6 * i + 2; }
    GL_HD u32 out_lo(u32 i) const { return 6 * i + 3; }
    GL_HD u32 out_hi(u32 i) const { return 6 * i + 4; }
    GL_HD u32 inverse(u32 i) const { return 6 * i + 5; }
    GL_HD u32 limb(u32 i, u32 j) const { return 6 * num_ops + 32 * i + j; }
};
struct U32AddManyLayout {   // per op: num_addends addends, carry in, result, carry out; then 16 + 2 two-bit limbs per op
    u32 num_addends, num_ops;
    GL_HD u32 addend(u32 i, u32 j) const { return (num_addends + 3) * i + j; }
    GL_HD u32 carry(u32 i) const { return (num_addends + 3) * i + num_addends; }
    GL_HD u32 result(u32 i) const { return (num_addends + 3) * i + num_addends + 1; }
    GL_HD u32 out_carry(u32 i) const { return (num_addends + 3) * i + num_addends + 2; }
    GL_HD u32 limb(u32 i, u32 j) const { return (num_addends + 3) * num_ops + 18 * i + j; }
};
struct U32SubLayout {   // per op: x, y, borrow in, result, borrow out; then 16 two-bit limbs per op
    u32 num_ops;
    GL_HD u32 x(u32 i) const { return 5 * i; }
    GL_HD u32 y(u32 i) const { return 5 * i + 1; }
    GL_HD u32 borrow(u32 i) const { return 5 * i + 2; }
    GL_HD u32 result(u32 i) const { return 5 * i + 3; }
    GL_HD u32 out_borrow(u32 i) const { return 5 * i + 4; }
    GL_HD u32 limb(u32 i, u32 j) const { return 5 * num_ops + 16 * i + j; }
};
struct ComparisonLayout {
    u32 num_bits, num_chunks;
    GL_HD u32 chunk_bits() const { return (num_bits + num_chunks - 1) / num_chunks; }
    GL_HD u32 first() const { return 0; }
    GL_HD u32 second() const { return 1; }
    GL_HD u32 result() const { return 2; }
    GL_HD u32 msd() const { return 3; }
    GL_HD u32 first_chunk(u32 i) const { return 4 + i; }
    GL_HD u32 second_chunk(u32 i) const { return 4 + num_chunks + i; }
    GL_HD u32 equality_dummy(u32 i) const { return 4 + 2 * num_chunks + i; }
    GL_HD u32 chunks_equal(u32 i) const { return 4 + 3 * num_chunks + i; }
    GL_HD u32 intermediate(u32 i) const { return 4 + 4 * num_chunks + i; }
    GL_HD u32 msd_bit(u32 i) const { return 4 + 5 * num_chunks + i; }
    GL_HD u32 num_wires() const { return 4 + 5 * num_chunks + chunk_bits() + 1; }
};
struct RandomAccessLayout {
    u32 bits, num_copies, num_extra;
    GL_HD u32 vec_size() const { return 1u << bits; }
    GL_HD u32 access_index(u32 c) const { return (2 + vec_size()) * c; }
    GL_HD u32 claimed(u32 c) const { return (2 + vec_size()) * c + 1; }
    GL_HD u32 item(u32 i, u32 c) const { return (2 + vec_size()) * c + 2 + i; }
    GL_HD u32 extra_const(u32 i) const { return (2 + vec_size()) * num_copies + i; }
    GL_HD u32 num_routed() const { return (2 + vec_size()) * num_copies + num_extra; }
    GL_HD u32 bit(u32 i, u32 c) const { return num_routed() + c * bits + i; }
};

struct InterleaveLayout {   // U32InterleaveGate: per op x, x_interleaved (routed), then 32 bits per op
    u32 num_ops;
    GL_HD u32 x(u32 i) const { return 2 * i; }
    GL_HD u32 x_interleaved(u32 i) const { return 2 * i + 1; }
    GL_HD u32 bit(u32 i, u32 j) const { return 2 * num_ops + 32 * i + j; }
};
struct UninterleaveLayout {   // UninterleaveToU32Gate / UninterleaveToB32Gate: per op x_interleaved, evens, odds (routed), then 64 bits per op
    u32 num_ops;
    GL_HD u32 x_interleaved(u32 i) const { return 3 * i; }
    GL_HD u32 evens(u32 i) const { return 3 * i + 1; }
    GL_HD u32 odds(u32 i) const { return 3 * i + 2; }
    GL_HD u32 bit(u32 i, u32 j) const { return 3 * num_ops + 64 * i + j; }
};
struct CosetInterpLayout {   // CosetInterpolationGate { subgroup_bits, degree }, D = 2 (recursive FRI verifier: interpolate_coset)
    u32 subgroup_bits, degree;
    GL_HD u32 num_points() const { return 1u << subgroup_bits; }
    GL_HD u32 num_intermediates() const { return (num_points() - 2) / (degree - 1); }
    GL_HD u32 shift() const { return 0; }
    GL_HD u32 value(u32 i) const { return 1 + 2 * i; }
    GL_HD u32 evaluation_point() const { return 1 + 2 * num_points(); }
    GL_HD u32 evaluation_value() const { return evaluation_point() + 2; }
    GL_HD u32 start_intermediates() const { return evaluation_value() + 2; }           // = number of routed wires
    GL_HD u32 intermediate_eval(u32 i) const { return start_intermediates() + 2 * i; }
    GL_HD u32 intermediate_prod(u32 i) const { return start_intermediates() + 2 * (num_intermediates() + i); }
    GL_HD u32 shifted_evaluation_point() const { return start_intermediates() + 4 * num_intermediates(); }
    GL_HD u32 num_wires() const { return shifted_evaluation_point() + 2; }
    // chunk c of the interpolation covers points [first(c), last(c)): degree points first, then degree - 1 per intermediate
    GL_HD u32 first(u32 c) const { return c == 0 ? 0 : 1 + (degree - 1) * c; }
    GL_HD u32 last(u32 c) const { u32 e = c == 0 ? degree : first(c) + degree - 1; return e < num_points() ? e : num_points(); }
};
// two_adic_subgroup(bits) and its barycentric weights w_i = 1 / prod_{j != i} (x_i - x_j)   (host)
static inline void plk_coset_domain(u32 bits, std::vector<u64> &domain, std::vector<u64> &weights) {
    const u32 n = 1u << bits;
    const u64 g = h_gl_root_of_unity((int)bits);
    domain.assign(n, 1);
    for (u32 i = 1; i < n; i++) domain[i] = h_gl_mul(domain[i - 1], g);
    weights.assign(n, 1);
    for (u32 i = 0; i < n; i++) {
        u64 d = 1;
        for (u32 j = 0; j < n; j++)
            if (j != i) d = h_gl_mul(d, gl_canon(gl_sub(domain[i], domain[j])));
        weights[i] = h_gl_inv(d);
    }
}

// MDS layer on builder values: out[r] = sum_i in[(i + r) % 12] * CIRC[i] + (r == 0) * 8 * in[0]
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BT>
GL_HD void plk_build_mds(BT &B, typename BT::V (&s)[12]) {
    typedef typename BT::V V;
    const u64 circ[12] = POSEIDON_MDS_CIRC_INIT;
    V o[12];
    for (int r = 0; r < 12; r++) {
        V acc = B.mul(s[r % 12], B.imm(circ[0]));
        for (int i = 1; i < 12; i++) acc = B.mad(s[(i + r) % 12], B.imm(circ[i]), acc);
        if (r == 0) acc = B.mad(s[0], B.imm(POSEIDON_MDS_DIAG0), acc);
        o[r] = acc;
    }
    for (int r = 0; r < 12; r++) s[r] = o[r];
}

// ---- the two gates whose programs need host-side tables: bytecode only (GvmBuilder) ----
static inline bool plk_build_poseidon_host(GvmBuilder &B, u32 num_wires) {
    typedef GvmVal V;
    const V one = B.imm(1);

        if (num_wires < 135) return false;
        V swap = B.wire(24);
        B.emit(B.mul(swap, B.sub(swap, one)));
        V st[12];
        for (u32 i = 0; i < 4; i++) B.emit(B.msub(swap, B.sub(B.wire(i + 4), B.wire(i)), B.wire(25 + i)));
        for (u32 i = 0; i < 4; i++) { st[i] = B.add(B.wire(i), B.wire(25 + i)); st[i + 4] = B.sub(B.wire(i + 4), B.wire(25 + i)); }
        for (u32 i = 8; i < 12; i++) st[i] = B.wire(i);
        u32 rnd = 0;
        for (u32 r = 0; r < 4; r++, rnd++) {
            for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_RC[12 * rnd + i]));
            if (r != 0)
                for (u32 i = 0; i < 12; i++) { V in = B.wire(29 + 12 * (r - 1) + i); B.emit(B.sub(st[i], in)); st[i] = in; }
            for (u32 i = 0; i < 12; i++) st[i] = B.pow(st[i], 7);
            plk_build_mds(B, st);
        }
        for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_FAST_FIRST[i]));
        {
            V o[11];
            for (u32 i = 0; i < 11; i++) {
                V acc = B.mul(st[1], B.imm(POSEIDON_FAST_INIT[11 * i]));
                for (u32 j = 1; j < 11; j++) acc = B.mad(st[j + 1], B.imm(POSEIDON_FAST_INIT[11 * i + j]), acc);
                o[i] = acc;
            }
            for (u32 i = 0; i < 11; i++) st[i + 1] = o[i];
        }
        for (u32 r = 0; r < 22; r++) {
            V in = B.wire(65 + r);
            B.emit(B.sub(st[0], in));
            V s0 = B.add(B.pow(in, 7), B.imm(POSEIDON_FAST_K[r]));
            V d = B.mul(s0, B.imm(25));
            for (u32 i = 0; i < 11; i++) d = B.mad(st[i + 1], B.imm(POSEIDON_FAST_ROW[11 * r + i]), d);
            for (u32 i = 0; i < 11; i++) st[i + 1] = B.mad(s0, B.imm(POSEIDON_FAST_COL[11 * r + i]), st[i + 1]);
            st[0] = d;
        }
        rnd += 22;
        for (u32 r = 0; r < 4; r++, rnd++) {
            for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_RC[12 * rnd + i]));
            for (u32 i = 0; i < 12; i++) { V in = B.wire(87 + 12 * r + i); B.emit(B.sub(st[i], in)); st[i] = in; }
            for (u32 i = 0; i < 12; i++) st[i] = B.pow(st[i], 7);
            plk_build_mds(B, st);
        }
        for (u32 i = 0; i < 12; i++) B.emit(B.sub(st[i], B.wire(12 + i)));
        return true;
    }
static inline bool plk_build_coset_host(GvmBuilder &B, const u32 p[4], u32 num_wires, u32 num_routed) {
    typedef GvmVal V;
    typedef GvmExtT<GvmBuilder> Ext;
    GvmExtOpsT<GvmBuilder> X(B);

        // p0 = subgroup_bits, p1 = degree.  shifted point * shift = evaluation point; barycentric interpolation of the 2^bits
        // values over the subgroup at the shifted point, cut into chunks whose running (eval, product) pairs are wires:
        //   eval' = eval * (z - x_i) + value_i * w_i * prod,   prod' = prod * (z - x_i)
        CosetInterpLayout L = {p[0], p[1]};
        if (L.subgroup_bits == 0 || L.subgroup_bits > 5 || L.degree < 2 || L.start_intermediates() > num_routed || L.num_wires() > num_wires) return false;
        std::vector<u64> domain, weights;
        plk_coset_domain(L.subgroup_bits, domain, weights);
        const V shift = B.wire(L.shift());
        const Ext z = X.wires(L.shifted_evaluation_point());
        X.emit(X.sub(X.wires(L.evaluation_point()), X.scale(z, shift)));
        Ext ev = {B.imm(0), B.imm(0)}, pr = {B.imm(1), B.imm(0)};
        for (u32 c = 0; c <= L.num_intermediates(); c++) {
            if (c > 0) {
                const Ext ie = X.wires(L.intermediate_eval(c - 1)), ip = X.wires(L.intermediate_prod(c - 1));
                X.emit(X.sub(ie, ev));
                X.emit(X.sub(ip, pr));
                ev = ie; pr = ip;
            }
            for (u32 i = L.first(c); i < L.last(c); i++) {
                const Ext term = {B.sub(z.a, B.imm(domain[i])), z.b};
                const Ext wv = X.scale(X.wires(L.value(i)), B.imm(weights[i]));
                ev = X.add(X.mul(ev, term), X.mul(wv, pr));
                pr = X.mul(pr, term);
            }
        }
        X.emit(X.sub(X.wires(L.evaluation_value()), ev));
        return true;
    }

// Appends the program of gate `kind` with parameters p[0..4) to the builder.  Returns false for an unknown kind or
// parameters that do not fit the wire / constant budget.
// BT is a BUILDER MODEL: GvmBuilder records bytecode (host); QuotDirect (plonk.cuh) evaluates the same formulas in place on
// the device -- one source for the interpreted and the natively compiled evaluators, so they cannot drift apart (and
// tests compare them gate by gate).  Gates whose programs need host-side tables (PoseidonGate: round-constant arrays;
// CosetInterpolationGate: barycentric weights) exist as bytecode only: in a direct model they return false.
#define PLK_MAX_LIST 64   // limbs / bits / items / chunks held at once by one gate
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BT>
GL_HD bool plk_build_gate(BT &B, u32 kind, const u32 p[4], u32 num_wires, u32 num_routed, u32 num_consts) {
    typedef typename BT::V V;
    typedef GvmExtT<BT> Ext;
    GvmExtOpsT<BT> X(B);
    const V one = B.imm(1);
    switch (kind) {
    case PLK_NOOP: return true;
    case PLK_CONSTANT: {   // const_i - wire_i
        if (p[0] > num_consts || p[0] > num_routed) return false;
        for (u32 i = 0; i < p[0]; i++) B.emit(B.sub(B.constant(i), B.wire(i)));
        return true;
    }
    case PLK_PUBLIC_INPUT: {   // wire_i - public_inputs_hash[i]
        for (u32 i = 0; i < 4; i++) B.emit(B.sub(B.wire(i), B.pi(i)));
        return true;
    }
    case PLK_ARITHMETIC: {   // out - (m0 * m1 * c0 + addend * c1)
        if (4 * p[0] > num_routed || num_consts < 2) return false;
        for (u32 i = 0; i < p[0]; i++) {
            V t = B.mul(B.mul(B.wire(4 * i), B.wire(4 * i + 1)), B.constant(0));
            V computed = B.mad(B.wire(4 * i + 2), B.constant(1), t);
            B.emit(B.sub(B.wire(4 * i + 3), computed));
        }
        return true;
    }
    case PLK_POSEIDON:
        if constexpr (BT::kDirect) return false;
        else return plk_build_poseidon_host(B, num_wires);
    case PLK_BASE_SUM: {   // p0 = B, p1 = num_limbs: sum_j limb_j B^j - sum, then prod_{k<B} (limb - k) per limb
        const u32 base = p[0], nl = p[1];
        if (base < 2 || nl == 0 || 1 + nl > num_routed) return false;
        if (nl > PLK_MAX_LIST) return false;
        V limbs[PLK_MAX_LIST];
        for (u32 j = 0; j < nl; j++) limbs[j] = B.wire(1 + j);
        B.emit(B.sub(B.reduce_with_powers(limbs, nl, base), B.wire(0)));
        for (u32 j = 0; j < nl; j++) B.emit(B.range_product(limbs[j], base));
        return true;
    }
    case PLK_ARITHMETIC_EXT: {   // per op (8 wires): output - (m0 * m1 * c0 + addend * c1) over F_p^2
        if (8 * p[0] > num_routed || num_consts < 2) return false;
        for (u32 i = 0; i < p[0]; i++) {
            Ext m0 = X.wires(8 * i), m1 = X.wires(8 * i + 2), ad = X.wires(8 * i + 4), out = X.wires(8 * i + 6);
            Ext computed = X.add(X.scale(X.mul(m0, m1), B.constant(0)), X.scale(ad, B.constant(1)));
            X.emit(X.sub(out, computed));
        }
        return true;
    }
    case PLK_MUL_EXT: {   // per op (6 wires): output - m0 * m1 * c0
        if (6 * p[0] > num_routed || num_consts < 1) return false;
        for (u32 i = 0; i < p[0]; i++) {
            Ext m0 = X.wires(6 * i), m1 = X.wires(6 * i + 2), out = X.wires(6 * i + 4);
            X.emit(X.sub(out, X.scale(X.mul(m0, m1), B.constant(0))));
        }
        return true;
    }
    case PLK_REDUCING: {   // output 0..2, alpha 2..4, old_acc 4..6, coeffs 6.., accs after; acc_i = acc_{i-1} * alpha + coeff_i
        const u32 nc = p[0];
        if (nc == 0 || 6 + nc > num_routed || 6 + nc + 2 * (nc - 1) > num_wires) return false;
        Ext alpha = X.wires(2), acc = X.wires(4);
        for (u32 i = 0; i < nc; i++) {
            Ext t = X.mul(acc, alpha);
            t.a = B.add(t.a, B.wire(6 + i));
            Ext nxt = i == nc - 1 ? X.wires(0) : X.wires(6 + nc + 2 * i);
            X.emit(X.sub(t, nxt));
            acc = nxt;
        }
        return true;
    }
    case PLK_REDUCING_EXT: {   // same with extension coefficients (2 wires each)
        const u32 nc = p[0];
        if (nc == 0 || 6 + 2 * nc > num_routed || 6 + 2 * nc + 2 * (nc - 1) > num_wires) return false;
        Ext alpha = X.wires(2), acc = X.wires(4);
        for (u32 i = 0; i < nc; i++) {
            Ext t = X.add(X.mul(acc, alpha), X.wires(6 + 2 * i));
            Ext nxt = i == nc - 1 ? X.wires(0) : X.wires(6 + 2 * nc + 2 * i);
            X.emit(X.sub(t, nxt));
            acc = nxt;
        }
        return true;
    }
    case PLK_RANDOM_ACCESS: {
        RandomAccessLayout L = {p[0], p[1], p[2]};
        if (L.bits == 0 || L.bits > 6 || L.num_routed() > num_routed || L.bit(L.bits - 1, L.num_copies - 1) >= num_wires || L.num_extra > num_consts) return false;
        for (u32 c = 0; c < L.num_copies; c++) {
            V bits[8], items[PLK_MAX_LIST];
            for (u32 i = 0; i < L.bits; i++) bits[i] = B.wire(L.bit(i, c));
            for (u32 i = 0; i < L.bits; i++) B.emit(B.mul(bits[i], B.sub(bits[i], one)));
            B.emit(B.sub(B.reduce_with_powers(bits, L.bits, 2), B.wire(L.access_index(c))));
            u32 count = L.vec_size();
            for (u32 i = 0; i < count; i++) items[i] = B.wire(L.item(i, c));
            for (u32 i = 0; i < L.bits; i++) {   // fold pairs with bit i: x + b * (y - x)   (in place: slot k / 2 <= k)
                for (u32 k = 0; k + 1 < count; k += 2) items[k / 2] = B.mad(bits[i], B.sub(items[k + 1], items[k]), items[k]);
                count >>= 1;
            }
            B.emit(B.sub(items[0], B.wire(L.claimed(c))));
        }
        for (u32 i = 0; i < L.num_extra; i++) B.emit(B.sub(B.constant(i), B.wire(L.extra_const(i))));
        return true;
    }
    case PLK_EXPONENTIATION: {   // base 0, power bits 1..1+n, output 1+n, intermediate values 2+n..2+2n
        const u32 n = p[0];
        if (n == 0 || 2 + 2 * n > num_wires || 2 + n > num_routed) return false;
        V base = B.wire(0);
        for (u32 i = 0; i < n; i++) {
            V prev = i == 0 ? one : B.mul(B.wire(2 + n + i - 1), B.wire(2 + n + i - 1));
            V bit = B.wire(1 + (n - 1 - i));
            // cur_bit * base + (1 - cur_bit)
            V sel = B.add(B.msub(bit, base, bit), one);
            B.emit(B.msub(prev, sel, B.wire(2 + n + i)));
        }
        B.emit(B.sub(B.wire(1 + n), B.wire(2 + n + n - 1)));
        return true;
    }
    case PLK_POSEIDON_MDS: {   // inputs 12 x 2 wires, outputs 12 x 2 wires: output_r - (MDS * inputs)_r
        if (48 > num_routed) return false;
        V lo[12], hi[12];
        for (u32 i = 0; i < 12; i++) { lo[i] = B.wire(2 * i); hi[i] = B.wire(2 * i + 1); }
        plk_build_mds(B, lo);
        plk_build_mds(B, hi);
        for (u32 r = 0; r < 12; r++) { B.emit(B.sub(B.wire(24 + 2 * r), lo[r])); B.emit(B.sub(B.wire(24 + 2 * r + 1), hi[r])); }
        return true;
    }
    case PLK_U32_ARITHMETIC: {
        U32ArithLayout L = {p[0]};
        if (p[0] == 0 || 6 * p[0] > num_routed || L.limb(p[0] - 1, 31) >= num_wires) return false;
        for (u32 i = 0; i < L.num_ops; i++) {
            V computed = B.mad(B.wire(L.m0(i)), B.wire(L.m1(i)), B.wire(L.addend(i)));
            V lo = B.wire(L.out_lo(i)), hi = B.wire(L.out_hi(i));
            V diff = B.sub(B.imm(0xFFFFFFFFull), hi);
            V hi_not_max = B.msub(B.wire(L.inverse(i)), diff, one);
            B.emit(B.mul(hi_not_max, lo));
            B.emit(B.sub(B.mad(hi, B.imm(1ull << 32), lo), computed));
            V clo = B.imm(0), chi = B.imm(0);
            bool have_lo = false, have_hi = false;
            for (int j = 31; j >= 0; j--) {
                V limb = B.wire(L.limb(i, (u32)j));
                B.emit(B.range_product(limb, 4));
                if (j < 16) { clo = have_lo ? B.mad(clo, B.imm(4), limb) : limb; have_lo = true; }
                else { chi = have_hi ? B.mad(chi, B.imm(4), limb) : limb; have_hi = true; }
            }
            B.emit(B.sub(clo, lo));
            B.emit(B.sub(chi, hi));
        }
        return true;
    }
    case PLK_U32_ADD_MANY: {
        U32AddManyLayout L = {p[0], p[1]};
        if (p[0] == 0 || p[1] == 0 || (p[0] + 3) * p[1] > num_routed || L.limb(p[1] - 1, 17) >= num_wires) return false;
        for (u32 i = 0; i < L.num_ops; i++) {
            V computed = B.wire(L.carry(i));
            for (u32 j = 0; j < L.num_addends; j++) computed = B.add(computed, B.wire(L.addend(i, j)));
            V res = B.wire(L.result(i)), oc = B.wire(L.out_carry(i));
            B.emit(B.sub(B.mad(oc, B.imm(1ull << 32), res), computed));
            V cres = B.imm(0), ccar = B.imm(0);
            bool have_r = false, have_c = false;
            for (int j = 17; j >= 0; j--) {
                V limb = B.wire(L.limb(i, (u32)j));
                B.emit(B.range_product(limb, 4));
                if (j < 16) { cres = have_r ? B.mad(cres, B.imm(4), limb) : limb; have_r = true; }
                else { ccar = have_c ? B.mad(ccar, B.imm(4), limb) : limb; have_c = true; }
            }
            B.emit(B.sub(cres, res));
            B.emit(B.sub(ccar, oc));
        }
        return true;
    }
    case PLK_U32_SUBTRACTION: {
        U32SubLayout L = {p[0]};
        if (p[0] == 0 || 5 * p[0] > num_routed || L.limb(p[0] - 1, 15) >= num_wires) return false;
        for (u32 i = 0; i < L.num_ops; i++) {
            V initial = B.sub(B.sub(B.wire(L.x(i)), B.wire(L.y(i))), B.wire(L.borrow(i)));
            V res = B.wire(L.result(i)), ob = B.wire(L.out_borrow(i));
            B.emit(B.sub(res, B.mad(ob, B.imm(1ull << 32), initial)));
            V comb = B.imm(0);
            bool have = false;
            for (int j = 15; j >= 0; j--) {
                V limb = B.wire(L.limb(i, (u32)j));
                B.emit(B.range_product(limb, 4));
                comb = have ? B.mad(comb, B.imm(4), limb) : limb; have = true;
            }
            B.emit(B.sub(comb, res));
            B.emit(B.mul(ob, B.sub(one, ob)));
        }
        return true;
    }
    case PLK_U32_RANGE_CHECK: {   // input limbs 0..k, 16 two-bit aux limbs each
        const u32 k = p[0];
        if (k == 0 || k > num_routed || k + 16 * k > num_wires) return false;
        for (u32 i = 0; i < k; i++) {
            V aux[16];
            for (u32 j = 0; j < 16; j++) aux[j] = B.wire(k + 16 * i + j);
            B.emit(B.sub(B.reduce_with_powers(aux, 16, 4), B.wire(i)));
            for (u32 j = 0; j < 16; j++) B.emit(B.range_product(aux[j], 4));
        }
        return true;
    }
    case PLK_COMPARISON: {
        ComparisonLayout L = {p[0], p[1]};
        if (p[0] == 0 || p[1] == 0 || p[1] > PLK_MAX_LIST || L.chunk_bits() > 4 || L.num_wires() > num_wires) return false;
        const u32 nch = L.num_chunks, cb = L.chunk_bits();
        V fc[PLK_MAX_LIST], sc[PLK_MAX_LIST];
        for (u32 i = 0; i < nch; i++) { fc[i] = B.wire(L.first_chunk(i)); sc[i] = B.wire(L.second_chunk(i)); }
        B.emit(B.sub(B.reduce_with_powers(fc, nch, 1ull << cb), B.wire(L.first())));
        B.emit(B.sub(B.reduce_with_powers(sc, nch, 1ull << cb), B.wire(L.second())));
        V msd_so_far = B.imm(0);
        for (u32 i = 0; i < nch; i++) {
            B.emit(B.range_product(fc[i], 1u << cb));
            B.emit(B.range_product(sc[i], 1u << cb));
            V diff = B.sub(sc[i], fc[i]);
            V eq = B.wire(L.chunks_equal(i));
            B.emit(B.sub(B.mul(diff, B.wire(L.equality_dummy(i))), B.sub(one, eq)));
            B.emit(B.mul(eq, diff));
            V inter = B.wire(L.intermediate(i));
            B.emit(B.sub(inter, B.mul(eq, msd_so_far)));
            msd_so_far = B.mad(B.sub(one, eq), diff, inter);
        }
        B.emit(B.sub(B.wire(L.msd()), msd_so_far));
        V bits[8];
        for (u32 i = 0; i <= cb; i++) bits[i] = B.wire(L.msd_bit(i));
        for (u32 i = 0; i <= cb; i++) B.emit(B.mul(bits[i], B.sub(one, bits[i])));
        B.emit(B.sub(B.add(B.imm(1ull << cb), B.wire(L.msd())), B.reduce_with_powers(bits, cb + 1, 2)));
        B.emit(B.sub(B.wire(L.result()), bits[cb]));
        return true;
    }
    case PLK_U32_INTERLEAVE: {   // bits boolean; x = sum b_j 2^j; x_interleaved = sum b_j 4^j (a zero bit between every two bits)
        InterleaveLayout L = {p[0]};
        if (p[0] == 0 || 2 * p[0] > num_routed || L.bit(p[0] - 1, 31) >= num_wires) return false;
        for (u32 i = 0; i < L.num_ops; i++) {
            V lin = B.imm(0), spread = B.imm(0);
            for (int j = 31; j >= 0; j--) {
                const V b = B.wire(L.bit(i, (u32)j));
                B.emit(B.mul(b, B.sub(one, b)));
                lin = j == 31 ? b : B.mad(lin, B.imm(2), b);
                spread = j == 31 ? b : B.mad(spread, B.imm(4), b);
            }
            B.emit(B.sub(B.wire(L.x(i)), lin));
            B.emit(B.sub(B.wire(L.x_interleaved(i)), spread));
        }
        return true;
    }
    case PLK_UNINTERLEAVE_TO_U32: case PLK_UNINTERLEAVE_TO_B32: {
        // 64 bits of x_interleaved; the even-position and the odd-position bits are packed densely (to U32: base 2) or
        // stay spread out (to B32: base 4)
        UninterleaveLayout L = {p[0]};
        if (p[0] == 0 || 3 * p[0] > num_routed || L.bit(p[0] - 1, 63) >= num_wires) return false;
        const u64 base = kind == PLK_UNINTERLEAVE_TO_U32 ? 2 : 4;
        for (u32 i = 0; i < L.num_ops; i++) {
            V all = B.imm(0), ev = B.imm(0), od = B.imm(0);
            for (int j = 63; j >= 0; j--) {
                const V b = B.wire(L.bit(i, (u32)j));
                B.emit(B.mul(b, B.sub(one, b)));
                all = j == 63 ? b : B.mad(all, B.imm(2), b);
                if (j & 1) od = j == 63 ? b : B.mad(od, B.imm(base), b);
                else ev = j == 62 ? b : B.mad(ev, B.imm(base), b);
            }
            B.emit(B.sub(B.wire(L.x_interleaved(i)), all));
            B.emit(B.sub(B.wire(L.evens(i)), ev));
            B.emit(B.sub(B.wire(L.odds(i)), od));
        }
        return true;
    }
    case PLK_COSET_INTERPOLATION:
        if constexpr (BT::kDirect) return false;
        else return plk_build_coset_host(B, p, num_wires, num_routed);
    }
    return false;
}
