// Gate library: bytecode programs (gate_vm.h) of the plonky2 / plonky2_crypto gates whose constraint formulas are restated
// in SURVEY.md A.8 and DESIGN.md 6.  Host code.  [DEP plonky2:gates/{constant,public_input,arithmetic_base,poseidon,base_sum,
// arithmetic_extension,multiplication_extension,reducing,reducing_extension,random_access,exponentiation,poseidon_mds,
// coset_interpolation}.rs]
// [DEP plonky2_crypto (plonky2_u32):gates/{arithmetic_u32,add_many_u32,subtraction_u32,range_check_u32,comparison}.rs]
// Tier C of SURVEY.md Appendix A: wire layouts and constraint ORDER are recalled, not checked against the absent source;
// with a Rust toolchain the programs are RECORDED from the gates' own eval_unfiltered_circuit instead and
// these builders become test fixtures.  Constraint t of a gate receives alpha^t, so order is part of the contract.
#pragma once
#include "gate_vm.h"
#include "poseidon_consts.h"

enum PlkGateKind : u32 {
    PLK_NOOP = 0, PLK_CONSTANT = 1, PLK_PUBLIC_INPUT = 2, PLK_ARITHMETIC = 3, PLK_POSEIDON = 4, PLK_BASE_SUM = 5,
    PLK_ARITHMETIC_EXT = 6, PLK_MUL_EXT = 7, PLK_REDUCING = 8, PLK_REDUCING_EXT = 9, PLK_RANDOM_ACCESS = 10,
    PLK_EXPONENTIATION = 11, PLK_POSEIDON_MDS = 12, PLK_U32_ARITHMETIC = 13, PLK_U32_ADD_MANY = 14, PLK_U32_SUBTRACTION = 15,
    PLK_U32_RANGE_CHECK = 16, PLK_COMPARISON = 17, PLK_COSET_INTERPOLATION = 18, PLK_NUM_KINDS = 19,
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
        case PLK_RANDOM_ACCESS: return p[0] + 1;
        case PLK_POSEIDON: return 7;
    }
    return 0;
}

// ---- wire layouts (shared with the synthetic witness filler) ----
struct U32ArithLayout {   // U32ArithmeticGate: 6 routed wires per op, then 32 two-bit limbs per op
    u32 num_ops;
    u32 m0(u32 i) const { return 6 * i; }
    u32 m1(u32 i) const { return 6 * i + 1; }
    u32 addend(u32 i) const { return 6 * i + 2; }
    u32 out_lo(u32 i) const { return 6 * i + 3; }
    u32 out_hi(u32 i) const { return 6 * i + 4; }
    u32 inverse(u32 i) const { return 6 * i + 5; }
    u32 limb(u32 i, u32 j) const { return 6 * num_ops + 32 * i + j; }
};
struct U32AddManyLayout {   // per op: num_addends addends, carry in, result, carry out; then 16 + 2 two-bit limbs per op
    u32 num_addends, num_ops;
    u32 addend(u32 i, u32 j) const { return (num_addends + 3) * i + j; }
    u32 carry(u32 i) const { return (num_addends + 3) * i + num_addends; }
    u32 result(u32 i) const { return (num_addends + 3) * i + num_addends + 1; }
    u32 out_carry(u32 i) const { return (num_addends + 3) * i + num_addends + 2; }
    u32 limb(u32 i, u32 j) const { return (num_addends + 3) * num_ops + 18 * i + j; }
};
struct U32SubLayout {   // per op: x, y, borrow in, result, borrow out; then 16 two-bit limbs per op
    u32 num_ops;
    u32 x(u32 i) const { return 5 * i; }
    u32 y(u32 i) const { return 5 * i + 1; }
    u32 borrow(u32 i) const { return 5 * i + 2; }
    u32 result(u32 i) const { return 5 * i + 3; }
    u32 out_borrow(u32 i) const { return 5 * i + 4; }
    u32 limb(u32 i, u32 j) const { return 5 * num_ops + 16 * i + j; }
};
struct ComparisonLayout {
    u32 num_bits, num_chunks;
    u32 chunk_bits() const { return (num_bits + num_chunks - 1) / num_chunks; }
    u32 first() const { return 0; }
    u32 second() const { return 1; }
    u32 result() const { return 2; }
    u32 msd() const { return 3; }
    u32 first_chunk(u32 i) const { return 4 + i; }
    u32 second_chunk(u32 i) const { return 4 + num_chunks + i; }
    u32 equality_dummy(u32 i) const { return 4 + 2 * num_chunks + i; }
    u32 chunks_equal(u32 i) const { return 4 + 3 * num_chunks + i; }
    u32 intermediate(u32 i) const { return 4 + 4 * num_chunks + i; }
    u32 msd_bit(u32 i) const { return 4 + 5 * num_chunks + i; }
    u32 num_wires() const { return 4 + 5 * num_chunks + chunk_bits() + 1; }
};
struct RandomAccessLayout {
    u32 bits, num_copies, num_extra;
    u32 vec_size() const { return 1u << bits; }
    u32 access_index(u32 c) const { return (2 + vec_size()) * c; }
    u32 claimed(u32 c) const { return (2 + vec_size()) * c + 1; }
    u32 item(u32 i, u32 c) const { return (2 + vec_size()) * c + 2 + i; }
    u32 extra_const(u32 i) const { return (2 + vec_size()) * num_copies + i; }
    u32 num_routed() const { return (2 + vec_size()) * num_copies + num_extra; }
    u32 bit(u32 i, u32 c) const { return num_routed() + c * bits + i; }
};

struct CosetInterpLayout {   // CosetInterpolationGate { subgroup_bits, degree }, D = 2 (recursive FRI verifier: interpolate_coset)
    u32 subgroup_bits, degree;
    u32 num_points() const { return 1u << subgroup_bits; }
    u32 num_intermediates() const { return (num_points() - 2) / (degree - 1); }
    u32 shift() const { return 0; }
    u32 value(u32 i) const { return 1 + 2 * i; }
    u32 evaluation_point() const { return 1 + 2 * num_points(); }
    u32 evaluation_value() const { return evaluation_point() + 2; }
    u32 start_intermediates() const { return evaluation_value() + 2; }           // = number of routed wires
    u32 intermediate_eval(u32 i) const { return start_intermediates() + 2 * i; }
    u32 intermediate_prod(u32 i) const { return start_intermediates() + 2 * (num_intermediates() + i); }
    u32 shifted_evaluation_point() const { return start_intermediates() + 4 * num_intermediates(); }
    u32 num_wires() const { return shifted_evaluation_point() + 2; }
    // chunk c of the interpolation covers points [first(c), last(c)): degree points first, then degree - 1 per intermediate
    u32 first(u32 c) const { return c == 0 ? 0 : 1 + (degree - 1) * c; }
    u32 last(u32 c) const { u32 e = c == 0 ? degree : first(c) + degree - 1; return e < num_points() ? e : num_points(); }
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
static inline void plk_build_mds(GvmBuilder &B, GvmVal (&s)[12]) {
    const u64 circ[12] = POSEIDON_MDS_CIRC_INIT;
    GvmVal o[12];
    for (int r = 0; r < 12; r++) {
        GvmVal acc = B.mul(s[r % 12], B.imm(circ[0]));
        for (int i = 1; i < 12; i++) acc = B.mad(s[(i + r) % 12], B.imm(circ[i]), acc);
        if (r == 0) acc = B.mad(s[0], B.imm(POSEIDON_MDS_DIAG0), acc);
        o[r] = acc;
    }
    for (int r = 0; r < 12; r++) s[r] = o[r];
}

// Appends the program of gate `kind` with parameters p[0..4) to the builder.  Returns false for an unknown kind or
// parameters that do not fit the wire / constant budget.
static inline bool plk_build_gate(GvmBuilder &B, u32 kind, const u32 p[4], u32 num_wires, u32 num_routed, u32 num_consts) {
    GvmExtOps X(B);
    const GvmVal one = B.imm(1);
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
            GvmVal t = B.mul(B.mul(B.wire(4 * i), B.wire(4 * i + 1)), B.constant(0));
            GvmVal computed = B.mad(B.wire(4 * i + 2), B.constant(1), t);
            B.emit(B.sub(B.wire(4 * i + 3), computed));
        }
        return true;
    }
    case PLK_POSEIDON: {
        if (num_wires < 135) return false;
        GvmVal swap = B.wire(24);
        B.emit(B.mul(swap, B.sub(swap, one)));
        GvmVal st[12];
        for (u32 i = 0; i < 4; i++) B.emit(B.msub(swap, B.sub(B.wire(i + 4), B.wire(i)), B.wire(25 + i)));
        for (u32 i = 0; i < 4; i++) { st[i] = B.add(B.wire(i), B.wire(25 + i)); st[i + 4] = B.sub(B.wire(i + 4), B.wire(25 + i)); }
        for (u32 i = 8; i < 12; i++) st[i] = B.wire(i);
        u32 rnd = 0;
        for (u32 r = 0; r < 4; r++, rnd++) {
            for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_RC[12 * rnd + i]));
            if (r != 0)
                for (u32 i = 0; i < 12; i++) { GvmVal in = B.wire(29 + 12 * (r - 1) + i); B.emit(B.sub(st[i], in)); st[i] = in; }
            for (u32 i = 0; i < 12; i++) st[i] = B.pow(st[i], 7);
            plk_build_mds(B, st);
        }
        for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_FAST_FIRST[i]));
        {
            GvmVal o[11];
            for (u32 i = 0; i < 11; i++) {
                GvmVal acc = B.mul(st[1], B.imm(POSEIDON_FAST_INIT[11 * i]));
                for (u32 j = 1; j < 11; j++) acc = B.mad(st[j + 1], B.imm(POSEIDON_FAST_INIT[11 * i + j]), acc);
                o[i] = acc;
            }
            for (u32 i = 0; i < 11; i++) st[i + 1] = o[i];
        }
        for (u32 r = 0; r < 22; r++) {
            GvmVal in = B.wire(65 + r);
            B.emit(B.sub(st[0], in));
            GvmVal s0 = B.add(B.pow(in, 7), B.imm(POSEIDON_FAST_K[r]));
            GvmVal d = B.mul(s0, B.imm(25));
            for (u32 i = 0; i < 11; i++) d = B.mad(st[i + 1], B.imm(POSEIDON_FAST_ROW[11 * r + i]), d);
            for (u32 i = 0; i < 11; i++) st[i + 1] = B.mad(s0, B.imm(POSEIDON_FAST_COL[11 * r + i]), st[i + 1]);
            st[0] = d;
        }
        rnd += 22;
        for (u32 r = 0; r < 4; r++, rnd++) {
            for (u32 i = 0; i < 12; i++) st[i] = B.add(st[i], B.imm(POSEIDON_RC[12 * rnd + i]));
            for (u32 i = 0; i < 12; i++) { GvmVal in = B.wire(87 + 12 * r + i); B.emit(B.sub(st[i], in)); st[i] = in; }
            for (u32 i = 0; i < 12; i++) st[i] = B.pow(st[i], 7);
            plk_build_mds(B, st);
        }
        for (u32 i = 0; i < 12; i++) B.emit(B.sub(st[i], B.wire(12 + i)));
        return true;
    }
    case PLK_BASE_SUM: {   // p0 = B, p1 = num_limbs: sum_j limb_j B^j - sum, then prod_{k<B} (limb - k) per limb
        const u32 base = p[0], nl = p[1];
        if (base < 2 || nl == 0 || 1 + nl > num_routed) return false;
        std::vector<GvmVal> limbs;
        for (u32 j = 0; j < nl; j++) limbs.push_back(B.wire(1 + j));
        B.emit(B.sub(B.reduce_with_powers(limbs, base), B.wire(0)));
        for (u32 j = 0; j < nl; j++) B.emit(B.range_product(limbs[j], base));
        return true;
    }
    case PLK_ARITHMETIC_EXT: {   // per op (8 wires): output - (m0 * m1 * c0 + addend * c1) over F_p^2
        if (8 * p[0] > num_routed || num_consts < 2) return false;
        for (u32 i = 0; i < p[0]; i++) {
            GvmExt m0 = X.wires(8 * i), m1 = X.wires(8 * i + 2), ad = X.wires(8 * i + 4), out = X.wires(8 * i + 6);
            GvmExt computed = X.add(X.scale(X.mul(m0, m1), B.constant(0)), X.scale(ad, B.constant(1)));
            X.emit(X.sub(out, computed));
        }
        return true;
    }
    case PLK_MUL_EXT: {   // per op (6 wires): output - m0 * m1 * c0
        if (6 * p[0] > num_routed || num_consts < 1) return false;
        for (u32 i = 0; i < p[0]; i++) {
            GvmExt m0 = X.wires(6 * i), m1 = X.wires(6 * i + 2), out = X.wires(6 * i + 4);
            X.emit(X.sub(out, X.scale(X.mul(m0, m1), B.constant(0))));
        }
        return true;
    }
    case PLK_REDUCING: {   // output 0..2, alpha 2..4, old_acc 4..6, coeffs 6.., accs after; acc_i = acc_{i-1} * alpha + coeff_i
        const u32 nc = p[0];
        if (nc == 0 || 6 + nc > num_routed || 6 + nc + 2 * (nc - 1) > num_wires) return false;
        GvmExt alpha = X.wires(2), acc = X.wires(4);
        for (u32 i = 0; i < nc; i++) {
            GvmExt t = X.mul(acc, alpha);
            t.a = B.add(t.a, B.wire(6 + i));
            GvmExt nxt = i == nc - 1 ? X.wires(0) : X.wires(6 + nc + 2 * i);
            X.emit(X.sub(t, nxt));
            acc = nxt;
        }
        return true;
    }
    case PLK_REDUCING_EXT: {   // same with extension coefficients (2 wires each)
        const u32 nc = p[0];
        if (nc == 0 || 6 + 2 * nc > num_routed || 6 + 2 * nc + 2 * (nc - 1) > num_wires) return false;
        GvmExt alpha = X.wires(2), acc = X.wires(4);
        for (u32 i = 0; i < nc; i++) {
            GvmExt t = X.add(X.mul(acc, alpha), X.wires(6 + 2 * i));
            GvmExt nxt = i == nc - 1 ? X.wires(0) : X.wires(6 + 2 * nc + 2 * i);
            X.emit(X.sub(t, nxt));
            acc = nxt;
        }
        return true;
    }
    case PLK_RANDOM_ACCESS: {
        RandomAccessLayout L = {p[0], p[1], p[2]};
        if (L.bits == 0 || L.bits > 6 || L.num_routed() > num_routed || L.bit(L.bits - 1, L.num_copies - 1) >= num_wires || L.num_extra > num_consts) return false;
        for (u32 c = 0; c < L.num_copies; c++) {
            std::vector<GvmVal> bits;
            for (u32 i = 0; i < L.bits; i++) bits.push_back(B.wire(L.bit(i, c)));
            for (u32 i = 0; i < L.bits; i++) B.emit(B.mul(bits[i], B.sub(bits[i], one)));
            B.emit(B.sub(B.reduce_with_powers(bits, 2), B.wire(L.access_index(c))));
            std::vector<GvmVal> items;
            for (u32 i = 0; i < L.vec_size(); i++) items.push_back(B.wire(L.item(i, c)));
            for (u32 i = 0; i < L.bits; i++) {   // fold pairs with bit i: x + b * (y - x)
                std::vector<GvmVal> nxt;
                for (size_t k = 0; k + 1 < items.size(); k += 2) nxt.push_back(B.mad(bits[i], B.sub(items[k + 1], items[k]), items[k]));
                items.swap(nxt);
            }
            B.emit(B.sub(items[0], B.wire(L.claimed(c))));
        }
        for (u32 i = 0; i < L.num_extra; i++) B.emit(B.sub(B.constant(i), B.wire(L.extra_const(i))));
        return true;
    }
    case PLK_EXPONENTIATION: {   // base 0, power bits 1..1+n, output 1+n, intermediate values 2+n..2+2n
        const u32 n = p[0];
        if (n == 0 || 2 + 2 * n > num_wires || 2 + n > num_routed) return false;
        GvmVal base = B.wire(0);
        for (u32 i = 0; i < n; i++) {
            GvmVal prev = i == 0 ? one : B.mul(B.wire(2 + n + i - 1), B.wire(2 + n + i - 1));
            GvmVal bit = B.wire(1 + (n - 1 - i));
            // cur_bit * base + (1 - cur_bit)
            GvmVal sel = B.add(B.msub(bit, base, bit), one);
            B.emit(B.msub(prev, sel, B.wire(2 + n + i)));
        }
        B.emit(B.sub(B.wire(1 + n), B.wire(2 + n + n - 1)));
        return true;
    }
    case PLK_POSEIDON_MDS: {   // inputs 12 x 2 wires, outputs 12 x 2 wires: output_r - (MDS * inputs)_r
        if (48 > num_routed) return false;
        GvmVal lo[12], hi[12];
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
            GvmVal computed = B.mad(B.wire(L.m0(i)), B.wire(L.m1(i)), B.wire(L.addend(i)));
            GvmVal lo = B.wire(L.out_lo(i)), hi = B.wire(L.out_hi(i));
            GvmVal diff = B.sub(B.imm(0xFFFFFFFFull), hi);
            GvmVal hi_not_max = B.msub(B.wire(L.inverse(i)), diff, one);
            B.emit(B.mul(hi_not_max, lo));
            B.emit(B.sub(B.mad(hi, B.imm(1ull << 32), lo), computed));
            GvmVal clo = B.imm(0), chi = B.imm(0);
            bool have_lo = false, have_hi = false;
            for (int j = 31; j >= 0; j--) {
                GvmVal limb = B.wire(L.limb(i, (u32)j));
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
            GvmVal computed = B.wire(L.carry(i));
            for (u32 j = 0; j < L.num_addends; j++) computed = B.add(computed, B.wire(L.addend(i, j)));
            GvmVal res = B.wire(L.result(i)), oc = B.wire(L.out_carry(i));
            B.emit(B.sub(B.mad(oc, B.imm(1ull << 32), res), computed));
            GvmVal cres = B.imm(0), ccar = B.imm(0);
            bool have_r = false, have_c = false;
            for (int j = 17; j >= 0; j--) {
                GvmVal limb = B.wire(L.limb(i, (u32)j));
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
            GvmVal initial = B.sub(B.sub(B.wire(L.x(i)), B.wire(L.y(i))), B.wire(L.borrow(i)));
            GvmVal res = B.wire(L.result(i)), ob = B.wire(L.out_borrow(i));
            B.emit(B.sub(res, B.mad(ob, B.imm(1ull << 32), initial)));
            GvmVal comb = B.imm(0);
            bool have = false;
            for (int j = 15; j >= 0; j--) {
                GvmVal limb = B.wire(L.limb(i, (u32)j));
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
            std::vector<GvmVal> aux;
            for (u32 j = 0; j < 16; j++) aux.push_back(B.wire(k + 16 * i + j));
            B.emit(B.sub(B.reduce_with_powers(aux, 4), B.wire(i)));
            for (u32 j = 0; j < 16; j++) B.emit(B.range_product(aux[j], 4));
        }
        return true;
    }
    case PLK_COMPARISON: {
        ComparisonLayout L = {p[0], p[1]};
        if (p[0] == 0 || p[1] == 0 || L.chunk_bits() > 4 || L.num_wires() > num_wires) return false;
        const u32 nch = L.num_chunks, cb = L.chunk_bits();
        std::vector<GvmVal> fc, sc;
        for (u32 i = 0; i < nch; i++) { fc.push_back(B.wire(L.first_chunk(i))); sc.push_back(B.wire(L.second_chunk(i))); }
        B.emit(B.sub(B.reduce_with_powers(fc, 1ull << cb), B.wire(L.first())));
        B.emit(B.sub(B.reduce_with_powers(sc, 1ull << cb), B.wire(L.second())));
        GvmVal msd_so_far = B.imm(0);
        for (u32 i = 0; i < nch; i++) {
            B.emit(B.range_product(fc[i], 1u << cb));
            B.emit(B.range_product(sc[i], 1u << cb));
            GvmVal diff = B.sub(sc[i], fc[i]);
            GvmVal eq = B.wire(L.chunks_equal(i));
            B.emit(B.sub(B.mul(diff, B.wire(L.equality_dummy(i))), B.sub(one, eq)));
            B.emit(B.mul(eq, diff));
            GvmVal inter = B.wire(L.intermediate(i));
            B.emit(B.sub(inter, B.mul(eq, msd_so_far)));
            msd_so_far = B.mad(B.sub(one, eq), diff, inter);
        }
        B.emit(B.sub(B.wire(L.msd()), msd_so_far));
        std::vector<GvmVal> bits;
        for (u32 i = 0; i <= cb; i++) bits.push_back(B.wire(L.msd_bit(i)));
        for (u32 i = 0; i <= cb; i++) B.emit(B.mul(bits[i], B.sub(one, bits[i])));
        B.emit(B.sub(B.add(B.imm(1ull << cb), B.wire(L.msd())), B.reduce_with_powers(bits, 2)));
        B.emit(B.sub(B.wire(L.result()), bits[cb]));
        return true;
    }
    case PLK_COSET_INTERPOLATION: {
        // p0 = subgroup_bits, p1 = degree.  shifted point * shift = evaluation point; barycentric interpolation of the 2^bits
        // values over the subgroup at the shifted point, cut into chunks whose running (eval, product) pairs are wires:
        //   eval' = eval * (z - x_i) + value_i * w_i * prod,   prod' = prod * (z - x_i)
        CosetInterpLayout L = {p[0], p[1]};
        if (L.subgroup_bits == 0 || L.subgroup_bits > 5 || L.degree < 2 || L.start_intermediates() > num_routed || L.num_wires() > num_wires) return false;
        std::vector<u64> domain, weights;
        plk_coset_domain(L.subgroup_bits, domain, weights);
        const GvmVal shift = B.wire(L.shift());
        const GvmExt z = X.wires(L.shifted_evaluation_point());
        X.emit(X.sub(X.wires(L.evaluation_point()), X.scale(z, shift)));
        GvmExt ev = {B.imm(0), B.imm(0)}, pr = {B.imm(1), B.imm(0)};
        for (u32 c = 0; c <= L.num_intermediates(); c++) {
            if (c > 0) {
                const GvmExt ie = X.wires(L.intermediate_eval(c - 1)), ip = X.wires(L.intermediate_prod(c - 1));
                X.emit(X.sub(ie, ev));
                X.emit(X.sub(ip, pr));
                ev = ie; pr = ip;
            }
            for (u32 i = L.first(c); i < L.last(c); i++) {
                const GvmExt term = {B.sub(z.a, B.imm(domain[i])), z.b};
                const GvmExt wv = X.scale(X.wires(L.value(i)), B.imm(weights[i]));
                ev = X.add(X.mul(ev, term), X.mul(wv, pr));
                pr = X.mul(pr, term);
            }
        }
        X.emit(X.sub(X.wires(L.evaluation_value()), ev));
        return true;
    }
    }
    return false;
}
