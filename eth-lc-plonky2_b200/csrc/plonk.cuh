// Rows a5 / a6: permutation-argument partial products and the quotient polynomials.
//
// Replaces plonky2::plonk::prover::{wires_permutation_partial_products_and_zs, compute_quotient_polys},
// plonky2::plonk::vanishing_poly::{eval_vanishing_poly_base_batch, evaluate_gate_constraints_base_batch},
// plonk_common::{check_partial_products, ZeroPolyOnCoset} and every Gate::eval_unfiltered_base_batch (dep plonky2 0.1.4,
// /root/reference/Cargo.lock:2347-2350; reached from /root/reference/eth-lc-plonky2/src/main.rs:230).  Gates are DATA:
// each gate of the circuit arrives as a bytecode program (gate_vm.h) -- recorded on the Rust side from the gate's own
// eval_unfiltered_circuit, or built by gate_lib.h for the gates restated there -- so the recursion gates
// (/root/reference/eth-lc-plonky2/src/targets.rs:468-470), plonky2_crypto's u32 gates (merkle_tree_gadget.rs:37) and
// BaseSumGate (utils.rs:102-103) need no per-gate port.  PoseidonGate additionally has a native evaluator (the FP64
// permutation of poseidon_f64.cuh), checked bit for bit against its bytecode.
//
// B200 design: one thread per point of the quotient domain, reading the column-major LDE of the three committed batches
// (coalesced).  The work is split into kernels that each fit the instruction cache and their register budget:
//   quot_perm_kernel      L_0(x)(Z - 1) and the partial-product checks            -> acc[c][pos]  =
//   quot_poseidon_kernel  PoseidonGate through the FP64 permutation                -> acc[c][pos] +=
//   quot_gates_kernel     every other gate through the bytecode interpreter        -> acc[c][pos] +=
//   quot_finish_kernel    * 1 / Z_H(x), scatter into natural order for the coset iNTT
// A gate streams sum_t alpha^(base + t) c_t and is multiplied by its selector filter ONCE.  L_0(x) comes from a per-circuit
// table built with a batched (Montgomery) inversion; the row-sequential Z accumulation of plonky2 is a three-phase
// multiplicative scan, with one batched inversion per row for the chunk denominators.
#pragma once
#include "gl64.cuh"
#include "gate_vm.h"
#include "gate_lib.h"
#include "poseidon.cuh"
#include "prover.cuh"

#define PLK_MAX_CHALLENGES 2
#define PLK_MAX_ROUTED 80
#define PLK_MAX_CHUNKS 40          // num_routed / quotient_degree_factor, quotient_degree_factor >= 2
#define PLK_MAX_ZH 16              // 2^quotient_degree_bits <= 2^rate_bits <= 16
#define PLK_UNUSED_SELECTOR 0xFFFFFFFFull

// ---- the alpha-weighted running sum of constraints for both challenges ----
// alpha^t comes from a table built on the host (every thread walks the same constraint sequence, so the loads are
// warp-uniform): one multiply-add per term and challenge.
// sum_t alpha^t term_t per challenge, LAZY: a 160-bit accumulator of unreduced 128-bit products (acc160, poseidon.cuh),
// reduced once per gate -- an emit is a product and a carry chain instead of a product and a reduction (2 of every 5
// field products of a gate-heavy circuit are these accumulations).
struct PlkAcc {
    acc160 s[PLK_MAX_CHALLENGES];
    const u64 *apow;               // [PLK_MAX_CHALLENGES][stride] powers of alpha (canonical)
    u32 stride;
    u32 t;                         // index of the next term
};
GL_HD void plk_acc_init(PlkAcc &a, const u64 *apow, u32 stride, u32 t0) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) a.s[c].lo = a.s[c].hi = 0, a.s[c].top = 0;
    a.apow = apow; a.stride = stride; a.t = t0;
}
GL_HD u64 plk_acc_value(const PlkAcc &a, int c) { return acc160_reduce(a.s[c]); }
GL_HD void plk_emit(PlkAcc &a, u64 term) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) acc160_mac(a.s[c], a.apow[c * a.stride + a.t], term);
    a.t++;
}
// the same for a term whose index is not the running one
GL_HD void plk_emit_at(PlkAcc &a, u32 t, u64 term) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) acc160_mac(a.s[c], a.apow[c * a.stride + t], term);
}
// The same interface with the sum reduced at every emit: 2 registers per challenge instead of 5.  For the evaluators
// that are short of registers and emit little (native PoseidonGate: 106 instead of 96 registers and spills with the lazy
// form; permutation checks: 22 terms).
struct PlkAccR {
    u64 sum[PLK_MAX_CHALLENGES];
    const u64 *apow;
    u32 stride;
    u32 t;
};
GL_HD void plk_acc_init(PlkAccR &a, const u64 *apow, u32 stride, u32 t0) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) a.sum[c] = 0;
    a.apow = apow; a.stride = stride; a.t = t0;
}
GL_HD u64 plk_acc_value(const PlkAccR &a, int c) { return a.sum[c]; }
GL_HD void plk_emit(PlkAccR &a, u64 term) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) a.sum[c] = gl_mul_add(a.apow[c * a.stride + a.t], term, a.sum[c]);
    a.t++;
}
GL_HD void plk_emit_at(PlkAccR &a, u32 t, u64 term) {
#pragma unroll
    for (int c = 0; c < PLK_MAX_CHALLENGES; c++) a.sum[c] = gl_mul_add(a.apow[c * a.stride + t], term, a.sum[c]);
}
// host: fills the table for the given challenges
static inline void plk_fill_apow(const u64 *alphas, u32 num_challenges, u32 stride, u64 *tab) {
    for (u32 c = 0; c < PLK_MAX_CHALLENGES; c++) {
        u64 a = c < num_challenges ? alphas[c] % GL_P : 0, v = 1;
        for (u32 t = 0; t < stride; t++) { tab[c * stride + t] = v; v = h_gl_mul(v, a); }
    }
}

// Accessor of the local wires / constants of one LDE point: column-major arrays with a row offset.
struct PlkCols {
    const u64 *base;
    u64 stride;   // elements between columns
    u64 row;
    GL_HD u64 operator[](u32 j) const { return base[(u64)j * stride + row]; }
};

// ---- native PoseidonGate evaluator: the constraints of [DEP plonky2:gates/poseidon.rs::eval_unfiltered_base_batch] with
// the linear layers of the permutation in exact FP64 (poseidon_f64.cuh).  The S-box inputs the gate constrains are the
// same values in the naive and in the "fast" partial rounds, so the constraint stream equals the bytecode's term for term
// (tests/test_gates_cpu.py, tests/test_gpu_plonk.py::test_native_poseidon_gate_equals_bytecode).
#ifndef PLK_F64_LANES
#define PLK_F64_LANES 3   // S-box lanes per rolled iteration (code size: the kernel must stay inside the instruction cache)
#endif
// the 22 partial rounds of the gate, two per step: pf_partial_rounds with both S-box inputs constrained and replaced
template <class W, class A>
GL_HD void plk_poseidon_gate_f64_partial(const W &w, A &acc, u64 (&st)[12]) {
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
        plk_emit(acc, gl_sub(a, in0));
        if (p != 0) {
#pragma unroll
            for (int j = 1; j < 12; j++) pf_renorm(al[j], ah[j]);
        }
        pf_pow7(in0, al[0], ah[0]);
        double t0l = PF_T(pair_t0)[p][0], t0h = PF_T(pair_t0)[p][1];
#pragma unroll
        for (int j = 0; j < 12; j++) {                  // row 0 of M = circ + 8 e_0 e_0^T
            t0l = pf_fma(al[j], PF_T(m_row0)[j], t0l);
            t0h = pf_fma(ah[j], PF_T(m_row0)[j], t0h);
        }
        const u64 b = pf_fold(t0l, t0h);
        const u64 in1 = w[65 + 2 * p + 1];
        plk_emit(acc, gl_sub(b, in1));
        double nl[12], nh[12];
        pf_circ12(al, PF_T(sc2), PF_T(pair_k_s)[p][0], nl);
        pf_circ12(ah, PF_T(sc2), PF_T(pair_k_s)[p][1], nh);
        double bl, bh;
        pf_pow7(in1, bl, bh);
        // lane 0 is REPLACED by in1^7: d = in1^7 - (computed b);  both rank-1 terms in one (pf_partial_rounds)
        const double ul = pf_fma(al[0], 8.0, bl - t0l), uh = pf_fma(ah[0], 8.0, bh - t0h);
#pragma unroll
        for (int r = 0; r < 12; r++) {
            al[r] = pf_fma(PF_T(c_col0)[r], ul, nl[r]);
            ah[r] = pf_fma(PF_T(c_col0)[r], uh, nh[r]);
        }
        al[0] = pf_fma(bl, 8.0, al[0]);
        ah[0] = pf_fma(bh, 8.0, ah[0]);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = pf_fold(al[i], ah[i]);          // state + RC_26
}
template <class W, class A>
GL_HD void plk_poseidon_gate_f64(const W &w, A &acc) {
    const u64 swap = w[24];
    plk_emit(acc, gl_mul(swap, gl_sub(swap, 1)));
    u64 st[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 lhs = w[i], rhs = w[i + 4], d = w[25 + i];
        plk_emit(acc, gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), d));
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
        if (L == 4) plk_poseidon_gate_f64_partial(w, acc, st);
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
                    plk_emit(acc, gl_sub(v, in));
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
        plk_emit(acc, gl_sub(st[0], w[12 + i]));
        u64 t = st[0];          // rotate by one so that the rolled loop always reads st[0]
#pragma unroll
        for (int k = 0; k < 11; k++) st[k] = st[k + 1];
        st[11] = t;
    }
}


struct PlkGateDev {            // one gate of the circuit as the kernels see it
    u32 prog_off, prog_len;    // words into QuotParams::prog
    u32 selector_index, group_start, group_end, row;   // row = index of the gate = the selector value that enables it
    u32 num_constraints;
    u32 native;                // 0: interpreted; 1: PoseidonGate through quot_poseidon_kernel; 2: compiled evaluator (quot_native_kernel)
    u32 kind, p[4];            // library kind and parameters (native == 2: what plk_build_gate<QuotDirect> is run with)
};

// compute_filter(row, group, s, many_selectors)
GL_HD u64 plk_filter(u32 row, u32 gs, u32 ge, u64 s, bool many) {
    u64 f = 1;
    for (u32 i = gs; i < ge; i++)
        if (i != row) f = gl_mul(f, gl_sub((u64)i, s));
    if (many) f = gl_mul(f, gl_sub(PLK_UNUSED_SELECTOR, s));
    return f;
}

struct QuotParams {
    u32 log_n, log_lq;             // rows; quotient domain = 2^log_lq = n * 2^quotient_degree_bits points (<= LDE size)
    u32 num_wires, num_routed, num_selectors, num_gate_constants, num_challenges, degree, npp;   // degree = quotient_degree_factor
    u64 stride;                    // elements between LDE columns (= LDE size L; = count for a row shard)
    const u64 *cs, *wires, *zs;    // LDE of constants||sigmas, wires, Z||partial products: [cols][stride], bit-reversed rows; the
                                   // quotient domain is the first 2^log_lq rows (natural LDE indices that are multiples of L / Lq)
    // Row shard (multi-GPU prover): the kernels cover positions [pos0, pos0 + count) of the quotient domain; row t of the
    // column arrays, of acc and (by_position) of out is position pos0 + t.  Single GPU: pos0 = 0, count = 2^log_lq.
    u64 pos0, count;
    u32 by_position;               // out[c * count + t] in position order instead of out[c * Lq + natural index]
    u64 k_is[PLK_MAX_ROUTED];
    u64 beta[PLK_MAX_CHALLENGES], gamma[PLK_MAX_CHALLENGES];
    const u64 *apow;               // plk_fill_apow(alpha): [PLK_MAX_CHALLENGES][apow_stride]
    u32 apow_stride, first_gate_term;
    const u64 *l0;                 // [Lq] by position: L_0(x) = Z_H(x) / (n (x - 1))
    u64 zh_inv[PLK_MAX_ZH];        // 1 / Z_H on the coset: index i mod 2^quotient_degree_bits
    const u64 *w_lo, *w_hi;        // two-level powers of w_Lq
    u32 w_lo_bits;
    u64 *acc;                      // [num_challenges][Lq] by position
    u64 *out;                      // [num_challenges][Lq], NATURAL order (input of the coset iNTT)
    const PlkGateDev *gates;       // device
    u32 num_gates;
    PlkGateDev poseidon;           // the natively evaluated gate (valid when has_poseidon)
    u32 has_poseidon;              // this launch sequence runs quot_poseidon_kernel: the interpreter skips native == 1 gates
    u32 use_native_gates;          // ... and quot_native_kernel: the interpreter skips native == 2 gates
    const u64 *prog, *imm;         // device: programs, immediates (imm[0..4) = public_inputs_hash)
};

GL_HD u64 quot_natural_index(const QuotParams &p, u64 pos) {
#ifdef __CUDA_ARCH__
    return p.log_lq ? (__brevll(pos) >> (64 - p.log_lq)) : 0;
#else
    u64 i = 0;
    for (u32 b = 0; b < p.log_lq; b++) i |= ((pos >> b) & 1) << (p.log_lq - 1 - b);
    return i;
#endif
}

// L_0(x)(Z - 1) and the partial-product checks: terms 0 .. first_gate_term of the alpha-sum.
GL_HD void quot_perm_point(const QuotParams &p, u64 t) {
    const u64 Lq = (u64)1 << p.log_lq;
    const u64 pos = p.pos0 + t;
    const u64 i = quot_natural_index(p, pos);
    const u64 i_next = (i + (Lq >> p.log_n)) & (Lq - 1);       // multiply x by w_n
    u64 pos_next;
    {
        QuotParams const &q = p;
#ifdef __CUDA_ARCH__
        pos_next = q.log_lq ? (__brevll(i_next) >> (64 - q.log_lq)) : 0;
#else
        pos_next = 0;
        for (u32 b = 0; b < q.log_lq; b++) pos_next |= ((i_next >> b) & 1) << (q.log_lq - 1 - b);
#endif
    }
    const u64 x = gl_mul(gl_mul(p.w_lo[i & ((1ull << p.w_lo_bits) - 1)], p.w_hi[i >> p.w_lo_bits]), 7);
    const u32 nc = p.num_selectors + p.num_gate_constants, nch = p.num_challenges, npp = p.npp;
    // i + Lq / n keeps the low bits of i, i.e. the high bits of pos: pos_next lies in the same row shard (<= 2^qdb shards)
    PlkCols sig = {p.cs + (u64)nc * p.stride, p.stride, t}, w = {p.wires, p.stride, t}, zs = {p.zs, p.stride, t}, zn = {p.zs, p.stride, pos_next - p.pos0};
    PlkAccR acc;
    plk_acc_init(acc, p.apow, p.apow_stride, 0);
    const u64 l0 = p.l0[pos];
    for (u32 c = 0; c < nch; c++) plk_emit(acc, gl_mul(l0, gl_sub(zs[c], 1)));
    // partial-product checks.  plonky2 orders the terms challenge-major (term index nch + c (npp + 1) + t); the loops run
    // chunk-major so that every wire and sigma value is loaded ONCE for all challenges (ncu, round 2: the challenge-major
    // loops read 23.0 GB for 13.5 GB of columns).  beta * k_j * x = (beta x) * k_j: one product per wire and challenge.
    u64 bx[PLK_MAX_CHALLENGES];
    for (u32 c = 0; c < nch; c++) bx[c] = gl_mul(p.beta[c], x);
    for (u32 t = 0; t <= npp; t++) {
        u64 num[PLK_MAX_CHALLENGES], den[PLK_MAX_CHALLENGES];
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) num[c] = den[c] = 1;
        for (u32 j = t * p.degree; j < (t + 1) * p.degree && j < p.num_routed; j++) {
            const u64 wv = w[j], sv = sig[j], kj = p.k_is[j];
#pragma unroll
            for (int c = 0; c < PLK_MAX_CHALLENGES; c++) {
                if ((u32)c < nch) {
                    num[c] = gl_mul(num[c], gl_add(gl_mul_add(bx[c], kj, wv), p.gamma[c]));
                    den[c] = gl_mul(den[c], gl_add(gl_mul_add(p.beta[c], sv, wv), p.gamma[c]));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) {
            if ((u32)c < nch) {
                const u64 prev = t == 0 ? zs[c] : zs[nch + c * npp + t - 1];
                const u64 next = t == npp ? zn[c] : zs[nch + c * npp + t];
                plk_emit_at(acc, nch + c * (npp + 1) + t, gl_sub(gl_mul(prev, num[c]), gl_mul(next, den[c])));
            }
        }
    }
    for (u32 c = 0; c < nch; c++) p.acc[(u64)c * p.count + t] = plk_acc_value(acc, c);
}

// PoseidonGate, native: acc += filter * sum_t alpha^(first_gate_term + t) c_t
GL_HD void quot_poseidon_point(const QuotParams &p, u64 t) {
    PlkCols cs = {p.cs, p.stride, t}, w = {p.wires, p.stride, t};
    const PlkGateDev &g = p.poseidon;
    const u64 filter = plk_filter(g.row, g.group_start, g.group_end, cs[g.selector_index], p.num_selectors > 1);
    PlkAccR acc;
    plk_acc_init(acc, p.apow, p.apow_stride, p.first_gate_term);
    plk_poseidon_gate_f64(w, acc);
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 *a = &p.acc[(u64)c * p.count + t];
        *a = gl_mul_add(filter, plk_acc_value(acc, c), *a);
    }
}

// Every interpreted gate.  plonky2 adds the gates' filtered constraints per constraint index and then takes powers of
// alpha; by linearity each gate streams c_t * alpha^(base + t) from the same starting power.
struct QuotGvmCtx {
    PlkCols w, k;          // wires; gate constants (after the selector prefix)
    const u64 *immv;
    GL_HD u64 wire(u32 i) const { return w[i]; }
    GL_HD u64 constant(u32 i) const { return k[i]; }
    GL_HD u64 imm(u32 i) const { return immv[i]; }
};
struct QuotGvmEmit {
    PlkAcc *acc;
    GL_HD void operator()(u64 v) { plk_emit(*acc, v); }
};
GL_HD void quot_gates_point(const QuotParams &p, u64 t) {
    PlkCols cs = {p.cs, p.stride, t};
    QuotGvmCtx cx = {{p.wires, p.stride, t}, {p.cs + (u64)p.num_selectors * p.stride, p.stride, t}, p.imm};
    u64 regs[GVM_NREG];
    u64 total[PLK_MAX_CHALLENGES] = {0, 0};
    for (u32 gi = 0; gi < p.num_gates; gi++) {
        const PlkGateDev g = p.gates[gi];
        if (g.prog_len == 0 || (g.native == 1 && p.has_poseidon) || (g.native == 2 && p.use_native_gates)) continue;
        const u64 filter = plk_filter(g.row, g.group_start, g.group_end, cs[g.selector_index], p.num_selectors > 1);
        PlkAcc acc;
        plk_acc_init(acc, p.apow, p.apow_stride, p.first_gate_term);
        QuotGvmEmit em = {&acc};
        gvm_run<GvmBaseField>(p.prog + g.prog_off, g.prog_len, cx, regs, em);
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) total[c] = gl_mul_add(filter, plk_acc_value(acc, c), total[c]);
    }
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 *a = &p.acc[(u64)c * p.count + t];
        *a = gl_add(*a, total[c]);
    }
}

// The compiled evaluators: gate_lib.h's formulas instantiated with a builder model that computes in place (V = u64).
// Same source as the bytecode, so the two cannot drift; tests compare them gate by gate (CPU replay and GPU).
struct QuotDirect {
    typedef u64 V;
    static constexpr bool kDirect = true;
    PlkCols w, k;
    const u64 *immv;
    PlkAcc *acc;
    GL_HD V wire(u32 i) const { return w[i]; }
    GL_HD V constant(u32 i) const { return k[i]; }
    GL_HD V pi(u32 i) const { return immv[i]; }
    GL_HD V imm(u64 v) const { return v; }
    GL_HD V add(V a, V b) const { return gl_add(a, b); }
    GL_HD V sub(V a, V b) const { return gl_sub(a, b); }
    GL_HD V mul(V a, V b) const { return gl_mul(a, b); }
    GL_HD V mad(V a, V b, V c) const { return gl_mul_add(a, b, c); }
    GL_HD V msub(V a, V b, V c) const { return gl_sub(gl_mul(a, b), c); }
    GL_HD void emit(V v) const { plk_emit(*acc, v); }
    GL_HD V pow(V x, u32 e) const {
        V r = 1, b = x;
        while (e) { if (e & 1) r = gl_mul(r, b); e >>= 1; if (e) b = gl_sqr(b); }
        return r;
    }
    GL_HD V reduce_with_powers(const V *v, u32 n, u64 base) const {
        V acc2 = v[n - 1];
        for (u32 i = n - 1; i-- > 0;) acc2 = gl_mul_add(acc2, base, v[i]);
        return acc2;
    }
    GL_HD V range_product(V x, u32 count) const {
        if (count == 4) {   // same identity as GvmBuilder::range_product
            const V y = gl_mul(x, gl_sub(x, 3));
            return gl_mul(y, gl_add(y, 2));
        }
        V acc2 = x;
        for (u32 j = 1; j < count; j++) acc2 = gl_mul(acc2, gl_sub(x, (u64)j));
        return acc2;
    }
};
// gates with native == 2 whose kind is in KIND_MASK: acc += filter * sum_t alpha^(first_gate_term + t) c_t.  The kind is
// dispatched through compile-time constants so that a kernel contains the evaluators of ITS kinds only.
template <u32 KIND_MASK>
GL_HD void quot_native_point(const QuotParams &p, u64 t) {
    const u32 kind_mask = KIND_MASK;
    PlkCols cs = {p.cs, p.stride, t};
    u64 total[PLK_MAX_CHALLENGES] = {0, 0};
    for (u32 gi = 0; gi < p.num_gates; gi++) {
        const PlkGateDev g = p.gates[gi];
        if (g.native != 2 || !p.use_native_gates || !((kind_mask >> g.kind) & 1)) continue;
        const u64 filter = plk_filter(g.row, g.group_start, g.group_end, cs[g.selector_index], p.num_selectors > 1);
        PlkAcc acc;
        plk_acc_init(acc, p.apow, p.apow_stride, p.first_gate_term);
        QuotDirect ev = {{p.wires, p.stride, t}, {p.cs + (u64)p.num_selectors * p.stride, p.stride, t}, p.imm, &acc};
#define PLK_NATIVE_CASE(K) if constexpr ((KIND_MASK >> (K)) & 1) { if (g.kind == (K)) plk_build_gate(ev, (u32)(K), g.p, p.num_wires, p.num_routed, p.num_gate_constants); }
        PLK_NATIVE_CASE(PLK_CONSTANT) PLK_NATIVE_CASE(PLK_PUBLIC_INPUT) PLK_NATIVE_CASE(PLK_ARITHMETIC) PLK_NATIVE_CASE(PLK_BASE_SUM)
        PLK_NATIVE_CASE(PLK_ARITHMETIC_EXT) PLK_NATIVE_CASE(PLK_MUL_EXT) PLK_NATIVE_CASE(PLK_REDUCING) PLK_NATIVE_CASE(PLK_REDUCING_EXT)
        PLK_NATIVE_CASE(PLK_RANDOM_ACCESS) PLK_NATIVE_CASE(PLK_EXPONENTIATION) PLK_NATIVE_CASE(PLK_POSEIDON_MDS)
        PLK_NATIVE_CASE(PLK_U32_ARITHMETIC) PLK_NATIVE_CASE(PLK_U32_ADD_MANY) PLK_NATIVE_CASE(PLK_U32_SUBTRACTION)
        PLK_NATIVE_CASE(PLK_U32_RANGE_CHECK) PLK_NATIVE_CASE(PLK_COMPARISON)
        PLK_NATIVE_CASE(PLK_U32_INTERLEAVE) PLK_NATIVE_CASE(PLK_UNINTERLEAVE_TO_U32) PLK_NATIVE_CASE(PLK_UNINTERLEAVE_TO_B32)
#undef PLK_NATIVE_CASE
#pragma unroll
        for (int c = 0; c < PLK_MAX_CHALLENGES; c++) total[c] = gl_mul_add(filter, plk_acc_value(acc, c), total[c]);
    }
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 *a = &p.acc[(u64)c * p.count + t];
        *a = gl_add(*a, total[c]);
    }
}
// the compiled evaluators are spread over three kernels by code size (arithmetic + recursion / lookups + exponentiation / u32)
#define PLK_NATIVE_GROUP_A ((1u << PLK_CONSTANT) | (1u << PLK_PUBLIC_INPUT) | (1u << PLK_ARITHMETIC) | (1u << PLK_BASE_SUM) | (1u << PLK_ARITHMETIC_EXT) | (1u << PLK_MUL_EXT) | (1u << PLK_REDUCING) | (1u << PLK_REDUCING_EXT))
#define PLK_NATIVE_GROUP_B ((1u << PLK_RANDOM_ACCESS) | (1u << PLK_EXPONENTIATION) | (1u << PLK_POSEIDON_MDS))
#define PLK_NATIVE_GROUP_C ((1u << PLK_U32_ARITHMETIC) | (1u << PLK_U32_ADD_MANY) | (1u << PLK_U32_SUBTRACTION) | (1u << PLK_U32_RANGE_CHECK) | (1u << PLK_COMPARISON) | \
                            (1u << PLK_U32_INTERLEAVE) | (1u << PLK_UNINTERLEAVE_TO_U32) | (1u << PLK_UNINTERLEAVE_TO_B32))
#define PLK_NATIVE_KINDS (PLK_NATIVE_GROUP_A | PLK_NATIVE_GROUP_B | PLK_NATIVE_GROUP_C)

// * 1 / Z_H(x); position -> natural index
GL_HD void quot_finish_point(const QuotParams &p, u64 t) {
    const u64 Lq = (u64)1 << p.log_lq;
    const u64 i = quot_natural_index(p, p.pos0 + t);
    const u64 zi = p.zh_inv[i & ((Lq >> p.log_n) - 1)];
    for (u32 c = 0; c < p.num_challenges; c++) {
        const u64 v = gl_canon(gl_mul(p.acc[(u64)c * p.count + t], zi));
        if (p.by_position) p.out[(u64)c * p.count + t] = v;
        else p.out[(u64)c * Lq + i] = v;
    }
}
// gathered row shards [G][num_challenges][Lq / G] in position order -> [num_challenges][Lq] in natural order
GL_HD void quot_unshard_point(const u64 *in, u64 *out, u32 log_lq, u32 log_shards, u32 nch, u64 pos) {
    const u64 Lq = (u64)1 << log_lq, count = Lq >> log_shards;
    const u64 g = pos / count, t = pos % count;
    u64 i = 0;
#ifdef __CUDA_ARCH__
    i = log_lq ? (__brevll(pos) >> (64 - log_lq)) : 0;
#else
    for (u32 b = 0; b < log_lq; b++) i |= ((pos >> b) & 1) << (log_lq - 1 - b);
#endif
    for (u32 c = 0; c < nch; c++) out[(u64)c * Lq + i] = in[(g * nch + c) * count + t];
}

// L_0 table: l0[pos] = zh[i mod 2^qdb] / (n (x_i - 1)), x_i = 7 w_Lq^i, i = bitrev(pos); 8 positions per thread share one
// inversion (Montgomery's trick).  x_i != 1 on the coset, so every factor is invertible.
struct L0Params {
    u32 log_n, log_lq;
    u64 n_field;
    u64 zh[PLK_MAX_ZH];
    const u64 *w_lo, *w_hi;
    u32 w_lo_bits;
    u64 *out;
};
#define L0_BATCH 8
GL_HD void l0_table_group(const L0Params &p, u64 grp) {
    const u64 Lq = (u64)1 << p.log_lq;
    u64 d[L0_BATCH], pre[L0_BATCH], idx[L0_BATCH];
    u64 run = 1;
    for (int k = 0; k < L0_BATCH; k++) {
        const u64 pos = grp * L0_BATCH + k;
        u64 i = 0;
        if (pos < Lq) {
#ifdef __CUDA_ARCH__
            i = p.log_lq ? (__brevll(pos) >> (64 - p.log_lq)) : 0;
#else
            for (u32 b = 0; b < p.log_lq; b++) i |= ((pos >> b) & 1) << (p.log_lq - 1 - b);
#endif
        }
        idx[k] = i;
        const u64 x = gl_mul(gl_mul(p.w_lo[i & ((1ull << p.w_lo_bits) - 1)], p.w_hi[i >> p.w_lo_bits]), 7);
        d[k] = gl_mul(p.n_field, gl_sub(x, 1));
        pre[k] = run;
        run = gl_mul(run, d[k]);
    }
    u64 inv = gl_inverse(run);
    for (int k = L0_BATCH - 1; k >= 0; k--) {
        const u64 pos = grp * L0_BATCH + k;
        const u64 dinv = gl_mul(inv, pre[k]);
        inv = gl_mul(inv, d[k]);
        if (pos < Lq) p.out[pos] = gl_canon(gl_mul(p.zh[idx[k] & ((Lq >> p.log_n) - 1)], dinv));
    }
}

// ---- a5: partial products ----
struct PpParams {
    u32 log_n, num_routed, num_challenges, degree;   // degree = quotient_degree_factor
    const u64 *wires;     // [num_wires][n] values on the subgroup
    const u64 *sigmas;    // [num_routed][n] values of the sigma polynomials
    u64 k_is[PLK_MAX_ROUTED];
    u64 beta[PLK_MAX_CHALLENGES], gamma[PLK_MAX_CHALLENGES];
    const u64 *w_lo, *w_hi;   // two-level powers of w_n
    u32 w_lo_bits;
    u64 *out;             // [num_challenges * (1 + npp)][n]: Z_0.., then per challenge the partial products
    u64 *row_prod;        // [num_challenges][n]: product of all chunk quotients of the row
};
// phase 1: per row, per challenge: the running products of the chunk quotients P_t = prod_{s<=t} q_s (t < npp) go to the
// partial-product columns, the full row product to row_prod.  The chunk denominators of a row and challenge are inverted
// together: one field inversion + 3 products per chunk (Montgomery's trick).
GL_HD void pp_row(const PpParams &p, u64 i) {
    const u64 n = (u64)1 << p.log_n;
    const u32 nchunks = (p.num_routed + p.degree - 1) / p.degree, npp = nchunks - 1;
    const u64 x = gl_mul(p.w_lo[i & ((1ull << p.w_lo_bits) - 1)], p.w_hi[i >> p.w_lo_bits]);
    for (u32 c = 0; c < p.num_challenges; c++) {
        u64 nums[PLK_MAX_CHUNKS], dens[PLK_MAX_CHUNKS], pre[PLK_MAX_CHUNKS];
        const u64 bx = gl_mul(p.beta[c], x);
        u64 run = 1;
        for (u32 t = 0; t < nchunks; t++) {
            u64 num = 1, den = 1;
            for (u32 j = t * p.degree; j < (t + 1) * p.degree && j < p.num_routed; j++) {
                u64 wv = p.wires[(u64)j * n + i];
                num = gl_mul(num, gl_add(gl_mul_add(bx, p.k_is[j], wv), p.gamma[c]));
                den = gl_mul(den, gl_add(gl_mul_add(p.beta[c], p.sigmas[(u64)j * n + i], wv), p.gamma[c]));
            }
            nums[t] = num; dens[t] = den; pre[t] = run;
            run = gl_mul(run, den);
        }
        u64 inv = gl_inverse(run);
        for (u32 t = nchunks; t-- > 0;) {   // nums[t] <- num_t / den_t
            nums[t] = gl_mul(nums[t], gl_mul(inv, pre[t]));
            inv = gl_mul(inv, dens[t]);
        }
        run = 1;
        for (u32 t = 0; t < nchunks; t++) {
            run = gl_mul(run, nums[t]);
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
__global__ void __launch_bounds__(128) quot_perm_kernel(QuotParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    quot_perm_point(p, t);
}
__global__ void __launch_bounds__(128) quot_poseidon_kernel(QuotParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    quot_poseidon_point(p, t);
}
__global__ void __launch_bounds__(128) quot_gates_kernel(QuotParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    quot_gates_point(p, t);
}
template <u32 KIND_MASK>
__global__ void __launch_bounds__(128) quot_native_kernel(QuotParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    quot_native_point<KIND_MASK>(p, t);
}
__global__ void __launch_bounds__(256) quot_unshard_kernel(const u64 *in, u64 *out, u32 log_lq, u32 log_shards, u32 nch) {
    const u64 pos = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >> log_lq) return;
    quot_unshard_point(in, out, log_lq, log_shards, nch, pos);
}
__global__ void __launch_bounds__(256) quot_finish_kernel(QuotParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    quot_finish_point(p, t);
}
__global__ void __launch_bounds__(128) l0_table_kernel(L0Params p) {
    const u64 grp = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if ((grp * L0_BATCH) >> p.log_lq) return;
    l0_table_group(p, grp);
}
// flag |= any non-zero element in [first, first + count) of each of `cols` columns (stride elements apart)
__global__ void __launch_bounds__(256) nonzero_flag_kernel(const u64 *data, u64 stride, u32 cols, u64 first, u64 count, u32 *flag) {
    const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    for (u32 c = 0; c < cols; c++)
        if (gl_canon(data[(u64)c * stride + first + k]) != 0) { *flag = 1; return; }
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
