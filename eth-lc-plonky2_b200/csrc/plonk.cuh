// Rows a5 / a6: permutation-argument partial products and the quotient polynomials.
//
// Replaces plonky2::plonk::prover::{wires_permutation_partial_products_and_zs, compute_quotient_polys},
// plonky2::plonk::vanishing_poly::{eval_vanishing_poly_base_batch, evaluate_gate_constraints_base_batch},
// plonk_common::{check_partial_products, ZeroPolyOnCoset} and Gate::eval_unfiltered_base_batch of the five core
// gates restated in SURVEY.md A.8 (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350; reached from
// /root/reference/eth-lc-plonky2/src/main.rs:230).  The other gates of the real eth-lc circuit (plonky2_crypto's u32 /
// comparison gates, the recursion gates) need their source and plug in as further cases of
// quot_gate_constraints.
//
// B200 design: one thread per LDE point, reading the column-major LDE of the three committed batches (coalesced);
// every gate's constraints are folded on the fly into sum_t alpha^t * filter * c_t for both challenges (no
// per-point constraint vectors, no re-packing into "batches of 32" as the CPU code does); the row-sequential Z
// accumulation of plonky2 becomes a three-phase multiplicative scan.
#pragma once
#include "gl64.cuh"
#include "poseidon.cuh"
#include "prover.cuh"

enum { PLK_NOOP = 0, PLK_CONSTANT = 1, PLK_PUBLIC_INPUT = 2, PLK_ARITHMETIC = 3, PLK_POSEIDON = 4 };
#define PLK_MAX_GATES 16
#define PLK_MAX_CHALLENGES 2
#define PLK_UNUSED_SELECTOR 0xFFFFFFFFull

struct PlkGate { u32 kind, selector_index, group_start, group_end; };
struct PlkCircuit {
    u32 degree_bits, num_wires, num_routed, num_gate_constants, num_selectors, num_challenges, quotient_degree_factor;
    u32 num_gates;
    PlkGate gates[PLK_MAX_GATES];
};

// ---- the alpha-weighted running sum of constraints for both challenges ----
// alpha^t comes from a table built on the host (every thread walks the same constraint sequence, so the loads are
// warp-uniform): one multiply-add per term and challenge instead of two products.
#define PLK_APOW_MAX 256
struct PlkAcc {
    u64 sum[PLK_MAX_CHALLENGES];   // sum_t alpha^t term_t
    const u64 *apow;               // [PLK_MAX_CHALLENGES][PLK_APOW_MAX] powers of alpha
    u32 t;                         // index of the next term
};
GL_HD void plk_emit(PlkAcc &a, u64 term) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) a.sum[c] = gl_mul_add(a.apow[c * PLK_APOW_MAX + a.t], term, a.sum[c]);
    a.t++;
}
GL_HD void plk_skip(PlkAcc &a, u32 k) { a.t += k; }  // advance over k absent constraints
// host: fills the table for the given challenges
static inline void plk_fill_apow(const u64 *alphas, u32 num_challenges, u64 *tab) {
    for (u32 c = 0; c < PLK_MAX_CHALLENGES; c++) {
        u64 a = c < num_challenges ? alphas[c] % GL_P : 0, v = 1;
        for (u32 t = 0; t < PLK_APOW_MAX; t++) { tab[c * PLK_APOW_MAX + t] = v; v = h_gl_mul(v, a); }
    }
}

// Accessor of the local wires / constants of one LDE point: column-major arrays with a row offset.
struct PlkCols {
    const u64 *base;
    u64 stride;   // elements between columns
    u64 row;
    GL_HD u64 operator[](u32 j) const { return base[(u64)j * stride + row]; }
};

// filter * constraints of one gate, streamed into `acc` starting at alpha^(first gate term)
template <class W>
GL_HD void plk_poseidon_gate(const W &w, u64 filter, PlkAcc &acc) {
    const u64 swap = w[24];
    plk_emit(acc, gl_mul(filter, gl_mul(swap, gl_sub(swap, 1))));
    u64 st[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 lhs = w[i], rhs = w[i + 4], d = w[25 + i];
        plk_emit(acc, gl_mul(filter, gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), d)));
        st[i] = gl_add(lhs, d);
        st[i + 4] = gl_sub(rhs, d);
    }
#pragma unroll
    for (int i = 8; i < 12; i++) st[i] = w[i];
    PSD_UNROLL1
    for (int r = 0; r < 4; r++) {
        PSD_UNROLL1
        for (int it = 0; it < 12 / PSD_SBOX_LANES; it++) {
#pragma unroll
            for (int k = 0; k < PSD_SBOX_LANES; k++) {
                int i = PSD_SBOX_LANES * it + k;
                u64 v = gl_add_c(st[k], PSD_RC(12 * r + i));
                if (r != 0) {
                    u64 in = w[29 + 12 * (r - 1) + i];
                    plk_emit(acc, gl_mul(filter, gl_sub(v, in)));
                    v = in;
                }
                st[k] = gl_pow7(v);
            }
            poseidon_rot(st);
        }
        poseidon_mds(st);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = gl_add_c(st[i], PSD_FIRST(i));
    {
        u64 o[11];
#pragma unroll
        for (int i = 0; i < 11; i++) o[i] = 0;
        PSD_UNROLL1
        for (int i = 0; i < 11; i++) {
            acc160 a = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < 11; j++) acc160_mac(a, PSD_INIT(11 * i + j), st[j + 1]);
            u64 v = acc160_reduce(a);
#pragma unroll
            for (int k = 0; k < 10; k++) o[k] = o[k + 1];
            o[10] = v;
        }
#pragma unroll
        for (int i = 0; i < 11; i++) st[i + 1] = o[i];
    }
    PSD_UNROLL1
    for (int r = 0; r < 22; r++) {
        u64 in = w[65 + r];
        plk_emit(acc, gl_mul(filter, gl_sub(st[0], in)));
        u64 s0 = gl_add_c(gl_pow7(in), PSD_K(r));
        acc160 a = {0, 0, 0};
        acc160_mac(a, s0, 25);
#pragma unroll
        for (int i = 0; i < 11; i++) acc160_mac(a, PSD_ROW(11 * r + i), st[i + 1]);
#pragma unroll
        for (int i = 0; i < 11; i++) st[i + 1] = gl_mul_add(PSD_COL(11 * r + i), s0, st[i + 1]);
        st[0] = acc160_reduce(a);
    }
    PSD_UNROLL1
    for (int r = 0; r < 4; r++) {
        PSD_UNROLL1
        for (int it = 0; it < 12 / PSD_SBOX_LANES; it++) {
#pragma unroll
            for (int k = 0; k < PSD_SBOX_LANES; k++) {
                int i = PSD_SBOX_LANES * it + k;
                u64 v = gl_add_c(st[k], PSD_RC(12 * (26 + r) + i));
                u64 in = w[87 + 12 * r + i];
                plk_emit(acc, gl_mul(filter, gl_sub(v, in)));
                st[k] = gl_pow7(in);
            }
            poseidon_rot(st);
        }
        poseidon_mds(st);
    }
    PSD_UNROLL1
    for (int i = 0; i < 12; i++) {
        plk_emit(acc, gl_mul(filter, gl_sub(st[0], w[12 + i])));
        // rotate by one so that the rolled loop always reads st[0]
        u64 t = st[0];
#pragma unroll
        for (int k = 0; k < 11; k++) st[k] = st[k + 1];
        st[11] = t;
    }
}

// The same constraints with the linear layers of the permutation in exact FP64 (poseidon_f64.cuh): the S-box inputs the
// gate constrains are the same values in the naive and in the "fast" partial rounds, so the constraint stream is
// identical term for term.  Default since the GPU parity run of round 1 (tests/test_gpu_plonk.py: 13 proofs bit-exact and
// verifying; quotient 35.1 -> 33.8 ms at 2^20 rows); -DPLK_POSEIDON_F64=0 keeps the integer evaluator for A/B.
#ifndef PLK_POSEIDON_F64
#define PLK_POSEIDON_F64 1
#endif
#ifndef PLK_F64_LANES
#define PLK_F64_LANES 3   // S-box lanes per rolled iteration (code size: the kernel must stay inside the instruction cache)
#endif
// the 22 partial rounds of the gate, two per step: pf_partial_rounds with both S-box inputs constrained and replaced
template <class W>
GL_HD void plk_poseidon_gate_f64_partial(const W &w, u64 filter, PlkAcc &acc, u64 (&st)[12]) {
    double al[12], ah[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        al[j] = pf_cvt((u32)st[j]);
        ah[j] = pf_cvt((u32)(st[j] >> 32));
    }
    al[0] -= 2251799813685248.0;
    ah[0] -= 2251799813685248.0;
    PSD_UNROLL1
    for (int p = 0; p < 11; p++) {
        const u64 a = pf_fold(al[0], ah[0]);
        const u64 in0 = w[65 + 2 * p];
        plk_emit(acc, gl_mul(filter, gl_sub(a, in0)));
#pragma unroll
        for (int j = 1; j < 12; j++) pf_renorm(al[j], ah[j]);
        pf_pow7(in0, al[0], ah[0]);
        double t0l = PF_T(pair_t0)[p][0], t0h = PF_T(pair_t0)[p][1];
#pragma unroll
        for (int j = 0; j < 12; j++) {
            t0l = pf_fma(al[j], PF_T(circ)[j], t0l);
            t0h = pf_fma(ah[j], PF_T(circ)[j], t0h);
        }
        t0l = pf_fma(al[0], 8.0, t0l);
        t0h = pf_fma(ah[0], 8.0, t0h);
        const u64 b = pf_fold(t0l, t0h);
        const u64 in1 = w[65 + 2 * p + 1];
        plk_emit(acc, gl_mul(filter, gl_sub(b, in1)));
        double nl[12], nh[12];
        pf_circ12(al, PF_T(sc2), PF_T(pair_k_s)[p][0], nl);
        pf_circ12(ah, PF_T(sc2), PF_T(pair_k_s)[p][1], nh);
#pragma unroll
        for (int r = 0; r < 12; r++) {
            nl[r] = pf_fma(al[0], PF_T(col8)[r], nl[r]);
            nh[r] = pf_fma(ah[0], PF_T(col8)[r], nh[r]);
        }
        nl[0] = pf_fma(t0l, 8.0, nl[0]);
        nh[0] = pf_fma(t0h, 8.0, nh[0]);
        double bl, bh;
        pf_pow7(in1, bl, bh);
        const double dl = bl - t0l, dh = bh - t0h;   // lane 0 is REPLACED by in1^7: the difference is against the computed b
#pragma unroll
        for (int r = 0; r < 12; r++) {
            al[r] = pf_fma(PF_T(m_col0)[r], dl, nl[r]);
            ah[r] = pf_fma(PF_T(m_col0)[r], dh, nh[r]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = pf_fold(al[i], ah[i]);          // state + RC_26
}
template <class W>
GL_HD void plk_poseidon_gate_f64(const W &w, u64 filter, PlkAcc &acc) {
    const u64 swap = w[24];
    plk_emit(acc, gl_mul(filter, gl_mul(swap, gl_sub(swap, 1))));
    u64 st[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 lhs = w[i], rhs = w[i + 4], d = w[25 + i];
        plk_emit(acc, gl_mul(filter, gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), d)));
        st[i] = gl_add(lhs, d);
        st[i + 4] = gl_sub(rhs, d);
    }
#pragma unroll
    for (int i = 8; i < 12; i++) st[i] = w[i];
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = gl_add_c(st[i], PSD_RC(i));     // state + RC_0
    // ONE rolled copy of the full round (layers 0..7; the partial rounds sit between 3 and 4): S-box on the wire values
    // (from round 1 on), MDS with the next round's constants in the chain heads, fold.  The lanes run PLK_F64_LANES per
    // iteration with the state rotated through the registers, as in poseidon_f64.cuh.
    PSD_UNROLL1
    for (int L = 0; L < 8; L++) {
        if (L == 4) plk_poseidon_gate_f64_partial(w, filter, acc, st);
        const u32 wire0 = L < 4 ? 29 + 12 * (L - 1) : 87 + 12 * (L - 4);   // sbox-in wires of this round (none for L = 0)
        double xl[12], xh[12];
#pragma unroll
        for (int k = 0; k < 12; k++) xl[k] = xh[k] = 0.0;
        PSD_UNROLL1
        for (int it = 0; it < 12 / PLK_F64_LANES; it++) {
            double tl[PLK_F64_LANES], th[PLK_F64_LANES];
#pragma unroll
            for (int k = 0; k < PLK_F64_LANES; k++) {
                u64 v = st[k];
                if (L != 0) {
                    const u64 in = w[wire0 + PLK_F64_LANES * it + k];
                    plk_emit(acc, gl_mul(filter, gl_sub(v, in)));
                    v = in;
                }
                pf_pow7(v, tl[k], th[k]);
            }
#pragma unroll
            for (int k = 0; k < 12 - PLK_F64_LANES; k++) {
                st[k] = st[k + PLK_F64_LANES];
                xl[k] = xl[k + PLK_F64_LANES];
                xh[k] = xh[k + PLK_F64_LANES];
            }
#pragma unroll
            for (int k = 0; k < PLK_F64_LANES; k++) {
                xl[12 - PLK_F64_LANES + k] = tl[k];
                xh[12 - PLK_F64_LANES + k] = th[k];
            }
        }
        double al[12], ah[12];
        pf_circ12(xl, PF_T(sc1), PF_T(full_init_s)[L][0], al);
        pf_circ12(xh, PF_T(sc1), PF_T(full_init_s)[L][1], ah);
        al[0] = pf_fma(xl[0], 8.0, al[0]);
        ah[0] = pf_fma(xh[0], 8.0, ah[0]);
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = pf_fold(al[i], ah[i]);
    }
    PSD_UNROLL1
    for (int i = 0; i < 12; i++) {
        plk_emit(acc, gl_mul(filter, gl_sub(st[0], w[12 + i])));
        u64 t = st[0];          // rotate by one so that the rolled loop always reads st[0]
#pragma unroll
        for (int k = 0; k < 11; k++) st[k] = st[k + 1];
        st[11] = t;
    }
}

// compute_filter(row, group, s, many_selectors)
GL_HD u64 plk_filter(u32 row, u32 gs, u32 ge, u64 s, bool many) {
    u64 f = 1;
    for (u32 i = gs; i < ge; i++)
        if (i != row) f = gl_mul(f, gl_sub((u64)i, s));
    if (many) f = gl_mul(f, gl_sub(PLK_UNUSED_SELECTOR, s));
    return f;
}

// All gate constraints of one point.  plonky2 adds the gates' filtered constraints per constraint index and then takes
// powers of alpha; by linearity each gate streams filter * c_t * alpha^(base + t) from the same starting power.
template <class W, class K>
GL_HD void plk_gate_constraints(const PlkCircuit &C, const W &w, const K &consts, const u64 *pi_hash, const PlkAcc &start, u64 *out) {
    u64 total[PLK_MAX_CHALLENGES] = {0, 0};
    for (u32 g = 0; g < C.num_gates; g++) {
        const PlkGate &gt = C.gates[g];
        if (gt.kind == PLK_NOOP) continue;
        u64 filter = plk_filter(g, gt.group_start, gt.group_end, consts[gt.selector_index], C.num_selectors > 1);
        PlkAcc acc = start;
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) acc.sum[c] = 0;
        if (gt.kind == PLK_CONSTANT) {
            for (u32 i = 0; i < 2; i++) plk_emit(acc, gl_mul(filter, gl_sub(consts[C.num_selectors + i], w[i])));
        } else if (gt.kind == PLK_PUBLIC_INPUT) {
            for (u32 i = 0; i < 4; i++) plk_emit(acc, gl_mul(filter, gl_sub(w[i], pi_hash[i])));
        } else if (gt.kind == PLK_ARITHMETIC) {
            const u64 c0 = consts[C.num_selectors], c1 = consts[C.num_selectors + 1];
            for (u32 i = 0; i < 20; i++) {
                u64 m0 = w[4 * i], m1 = w[4 * i + 1], ad = w[4 * i + 2], o = w[4 * i + 3];
                u64 computed = gl_add(gl_mul(gl_mul(m0, m1), c0), gl_mul(ad, c1));
                plk_emit(acc, gl_mul(filter, gl_sub(o, computed)));
            }
        } else if (gt.kind == PLK_POSEIDON) {
#if PLK_POSEIDON_F64
            plk_poseidon_gate_f64(w, filter, acc);
#else
            plk_poseidon_gate(w, filter, acc);
#endif
        }
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) total[c] = gl_add(total[c], acc.sum[c]);
    }
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) out[c] = total[c];
}

struct QuotParams {
    PlkCircuit C;
    u32 log_l;                     // degree_bits + 3
    const u64 *cs, *wires, *zs;    // LDE of constants||sigmas, wires, Z||partial products: [cols][L], bit-reversed rows
    u64 k_is[80];
    u64 beta[PLK_MAX_CHALLENGES], gamma[PLK_MAX_CHALLENGES], alpha[PLK_MAX_CHALLENGES];
    const u64 *apow;               // plk_fill_apow(alpha): [PLK_MAX_CHALLENGES][PLK_APOW_MAX]
    u64 pi_hash[4];
    u64 zh[8], zh_inv[8];          // ZeroPolyOnCoset: 7^n w_8^i - 1 and inverses
    u64 n_field;                   // n as a field element
    const u64 *w_lo, *w_hi;        // two-level powers of w_L
    u32 w_lo_bits;
    u64 *out;                      // [num_challenges][L], NATURAL order (input of the coset iNTT)
};

// One LDE point: position `pos` of the bit-reversed storage, natural index i = bitrev(pos).
GL_HD void quot_point(const QuotParams &p, u64 pos) {
    const PlkCircuit &C = p.C;
    const u64 L = (u64)1 << p.log_l;
    u64 i = 0;
    for (u32 b = 0; b < p.log_l; b++) i |= ((pos >> b) & 1) << (p.log_l - 1 - b);
    const u64 i_next = (i + 8) & (L - 1);
    u64 pos_next = 0;
    for (u32 b = 0; b < p.log_l; b++) pos_next |= ((i_next >> b) & 1) << (p.log_l - 1 - b);
    const u64 x = gl_mul(gl_mul(p.w_lo[i & ((1ull << p.w_lo_bits) - 1)], p.w_hi[i >> p.w_lo_bits]), 7);
    const u32 nc = C.num_selectors + C.num_gate_constants, nch = C.num_challenges;
    const u32 npp = (C.num_routed + C.quotient_degree_factor - 1) / C.quotient_degree_factor - 1;
    PlkCols cs = {p.cs, L, pos}, sig = {p.cs + (u64)nc * L, L, pos}, w = {p.wires, L, pos}, zs = {p.zs, L, pos}, zn = {p.zs, L, pos_next};
    PlkAcc acc;
    for (u32 c = 0; c < PLK_MAX_CHALLENGES; c++) acc.sum[c] = 0;
    acc.apow = p.apow; acc.t = 0;
    // L_0(x) (Z - 1)
    const u64 l0 = gl_mul(p.zh[i & 7], gl_inverse(gl_mul(p.n_field, gl_sub(x, 1))));
    for (u32 c = 0; c < nch; c++) plk_emit(acc, gl_mul(l0, gl_sub(zs[c], 1)));
    // partial-product checks, challenge-major.  beta * k_j * x = (beta x) * k_j: one product per wire and challenge
    for (u32 c = 0; c < nch; c++) {
        const u64 bx = gl_mul(p.beta[c], x);
        for (u32 t = 0; t <= npp; t++) {
            u64 prev = t == 0 ? zs[c] : zs[nch + c * npp + t - 1];
            u64 next = t == npp ? zn[c] : zs[nch + c * npp + t];
            u64 num = 1, den = 1;
            for (u32 j = t * C.quotient_degree_factor; j < (t + 1) * C.quotient_degree_factor && j < C.num_routed; j++) {
                u64 wv = w[j];
                num = gl_mul(num, gl_add(gl_mul_add(bx, p.k_is[j], wv), p.gamma[c]));
                den = gl_mul(den, gl_add(gl_mul_add(p.beta[c], sig[j], wv), p.gamma[c]));
            }
            plk_emit(acc, gl_sub(gl_mul(prev, num), gl_mul(next, den)));
        }
    }
    u64 gate_sum[PLK_MAX_CHALLENGES];
    plk_gate_constraints(C, w, cs, p.pi_hash, acc, gate_sum);
    for (u32 c = 0; c < nch; c++) {
        u64 v = gl_mul(gl_add(acc.sum[c], gate_sum[c]), p.zh_inv[i & 7]);
        p.out[(u64)c * L + i] = gl_canon(v);
    }
}

// ---- a5: partial products ----
struct PpParams {
    u32 log_n, num_routed, num_challenges, degree;   // degree = quotient_degree_factor
    const u64 *wires;     // [num_wires][n] values on the subgroup
    const u64 *sigmas;    // [num_routed][n] values of the sigma polynomials
    u64 k_is[80];
    u64 beta[PLK_MAX_CHALLENGES], gamma[PLK_MAX_CHALLENGES];
    const u64 *w_lo, *w_hi;   // two-level powers of w_n
    u32 w_lo_bits;
    u64 *out;             // [num_challenges * (1 + npp)][n]: Z_0.., then per challenge the partial products
    u64 *row_prod;        // [num_challenges][n]: product of all chunk quotients of the row
};
// phase 1: per row, per challenge: the running products of the chunk quotients P_t = prod_{s<=t} q_s (t < npp) go to the
// partial-product columns, the full row product to row_prod.
GL_HD void pp_row(const PpParams &p, u64 i) {
    const u64 n = (u64)1 << p.log_n;
    const u32 nchunks = (p.num_routed + p.degree - 1) / p.degree, npp = nchunks - 1;
    const u64 x = gl_mul(p.w_lo[i & ((1ull << p.w_lo_bits) - 1)], p.w_hi[i >> p.w_lo_bits]);
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 run = 1;
        const u64 bx = gl_mul(p.beta[c], x);
        for (u32 t = 0; t < nchunks; t++) {
            u64 num = 1, den = 1;
            for (u32 j = t * p.degree; j < (t + 1) * p.degree && j < p.num_routed; j++) {
                u64 wv = p.wires[(u64)j * n + i];
                num = gl_mul(num, gl_add(gl_mul_add(bx, p.k_is[j], wv), p.gamma[c]));
                den = gl_mul(den, gl_add(gl_mul_add(p.beta[c], p.sigmas[(u64)j * n + i], wv), p.gamma[c]));
            }
            run = gl_mul(run, gl_mul(num, gl_inverse(den)));
            if (t < npp) p.out[((u64)p.num_challenges + (u64)c * npp + t) * n + i] = run;
        }
        p.row_prod[(u64)c * n + i] = run;
    }
}
// phase 3 (after the exclusive scan of row_prod into the Z columns): partial products *= Z(x_i)
GL_HD void pp_finish(const PpParams &p, u64 i) {
    const u64 n = (u64)1 << p.log_n;
    const u32 npp = (p.num_routed + p.degree - 1) / p.degree - 1;
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 z = p.out[(u64)c * n + i];
        for (u32 t = 0; t < npp; t++) {
            u64 *q = &p.out[((u64)p.num_challenges + (u64)c * npp + t) * n + i];
            *q = gl_canon(gl_mul(*q, z));
        }
    }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128) quotient_kernel(QuotParams p) {
    const u64 pos = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >> p.log_l) return;
    quot_point(p, pos);
}
__global__ void __launch_bounds__(128) pp_row_kernel(PpParams p) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> p.log_n) return;
    pp_row(p, i);
}
__global__ void __launch_bounds__(256) pp_finish_kernel(PpParams p) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> p.log_n) return;
    pp_finish(p, i);
}
// exclusive multiplicative scan over rows, three phases.  SCAN_B elements per block.
#define SCAN_B 1024
__global__ void __launch_bounds__(256) scan_block_prod_kernel(const u64 *in, u64 n, u64 *block_prod) {
    __shared__ u64 sm[256];
    const u64 base = (u64)blockIdx.x * SCAN_B;
    u64 v = 1;
    for (u32 k = 0; k < SCAN_B / 256; k++) {
        u64 idx = base + threadIdx.x * (SCAN_B / 256) + k;
        if (idx < n) v = gl_mul(v, in[idx]);
    }
    sm[threadIdx.x] = v;
    __syncthreads();
    for (u32 h = 128; h >= 1; h >>= 1) {
        if (threadIdx.x < h) sm[threadIdx.x] = gl_mul(sm[threadIdx.x], sm[threadIdx.x + h]);
        __syncthreads();
    }
    if (threadIdx.x == 0) block_prod[blockIdx.x] = sm[0];
}
__global__ void scan_block_offsets_kernel(u64 *block_prod, u64 nb) {   // single thread block, sequential over <= 2^14 blocks
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        u64 acc = 1;
        for (u64 b = 0; b < nb; b++) { u64 v = block_prod[b]; block_prod[b] = acc; acc = gl_mul(acc, v); }
    }
}
__global__ void __launch_bounds__(256) scan_apply_kernel(const u64 *in, u64 n, const u64 *block_off, u64 *out) {
    __shared__ u64 sm[256];
    const u64 base = (u64)blockIdx.x * SCAN_B;
    const u32 per = SCAN_B / 256;
    u64 loc[SCAN_B / 256];
    u64 v = 1;
    for (u32 k = 0; k < per; k++) {
        u64 idx = base + threadIdx.x * per + k;
        loc[k] = idx < n ? in[idx] : 1;
        v = gl_mul(v, loc[k]);
    }
    sm[threadIdx.x] = v;
    __syncthreads();
    // exclusive prefix over the 256 thread products (Hillis-Steele on a copy)
    for (u32 d = 1; d < 256; d <<= 1) {
        u64 t = threadIdx.x >= d ? sm[threadIdx.x - d] : 1;
        __syncthreads();
        sm[threadIdx.x] = gl_mul(sm[threadIdx.x], t);
        __syncthreads();
    }
    u64 acc = gl_mul(block_off[blockIdx.x], threadIdx.x ? sm[threadIdx.x - 1] : 1);
    for (u32 k = 0; k < per; k++) {
        u64 idx = base + threadIdx.x * per + k;
        if (idx < n) out[idx] = gl_canon(acc);
        acc = gl_mul(acc, loc[k]);
    }
}
__global__ void __launch_bounds__(256) coset_unscale_kernel(u64 *coeffs, u64 n, u32 cols, const u64 *s_lo, const u64 *s_hi, u32 lo_bits) {
    const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const u64 s = gl_mul(s_lo[k & ((1ull << lo_bits) - 1)], s_hi[k >> lo_bits]);   // 7^-k
    for (u32 c = 0; c < cols; c++) coeffs[(u64)c * n + k] = gl_canon(gl_mul(coeffs[(u64)c * n + k], s));
}
#endif
