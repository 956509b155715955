// Poseidon over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds), one sponge state per thread.
//
// Replaces plonky2::hash::poseidon::Poseidon::poseidon for GoldilocksField and the sponge helpers of
// plonky2::hash::hashing (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350).
// Same outputs as the reference ("naive") round function; the partial rounds run in the algebraically
// identical sparse form (constants derived by tools/gen_poseidon_consts.py and checked against the
// upstream known-answer vectors).
//
// Code-size discipline (profiles/r01_leaves_v0.md): a fully unrolled permutation is ~14 K instructions (217 KB) and
// the kernel stalled on instruction fetch 13 of every 14 issue slots.  Here every loop is rolled and the
// per-lane work of a full round runs 3 lanes per iteration with the state ROTATED through the register file
// (no dynamic register indexing exists), so the hot loops together stay inside the 32 KB instruction cache.
//
// Pipe budget per permutation: S-boxes 118 * 4 field products; full-round MDS as
// 2 * 144 IMAD.WIDE on 32-bit halves (sums < 2^42, one cheap fold per lane); partial rounds as 22 * (11
// products into a 160-bit accumulator + 11 multiply-adds).
#pragma once
#include "gl64.cuh"
#include "poseidon_consts.h"

#ifdef __CUDACC__
__constant__ u64 c_rc[360];
__constant__ u64 c_fast_first[12];
__constant__ u64 c_fast_k[22];
__constant__ u64 c_fast_row[22 * 11];
__constant__ u64 c_fast_col[22 * 11];
__constant__ u64 c_fast_init[11 * 11];
__constant__ u32 c_mds[13] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20, 8};  // CIRC[0..12), DIAG[0]
#endif
#ifdef __CUDA_ARCH__
#define PSD_RC(i) c_rc[i]
#define PSD_FIRST(i) c_fast_first[i]
#define PSD_K(i) c_fast_k[i]
#define PSD_ROW(i) c_fast_row[i]
#define PSD_COL(i) c_fast_col[i]
#define PSD_INIT(i) c_fast_init[i]
#define PSD_UNROLL1 _Pragma("unroll 1")
#else
#define PSD_RC(i) POSEIDON_RC[i]
#define PSD_FIRST(i) POSEIDON_FAST_FIRST[i]
#define PSD_K(i) POSEIDON_FAST_K[i]
#define PSD_ROW(i) POSEIDON_FAST_ROW[i]
#define PSD_COL(i) POSEIDON_FAST_COL[i]
#define PSD_INIT(i) POSEIDON_FAST_INIT[i]
#define PSD_UNROLL1
#endif

#define POSEIDON_RC_AT(i) PSD_RC(i)
#include "poseidon_f64.cuh"

#ifdef __CUDACC__
// Uploads the constant tables of this translation unit; call once per module before the first launch.
static inline cudaError_t poseidon_upload_constants() {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_rc, POSEIDON_RC, sizeof(POSEIDON_RC))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_first, POSEIDON_FAST_FIRST, sizeof(POSEIDON_FAST_FIRST))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_k, POSEIDON_FAST_K, sizeof(POSEIDON_FAST_K))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_row, POSEIDON_FAST_ROW, sizeof(POSEIDON_FAST_ROW))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_col, POSEIDON_FAST_COL, sizeof(POSEIDON_FAST_COL))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_init, POSEIDON_FAST_INIT, sizeof(POSEIDON_FAST_INIT))) != cudaSuccess) return e;
    return psd_f64_upload_tables();
}
#endif

// 160-bit accumulator for sums of up to 2^32 full 128-bit products.
struct acc160 {
    u64 lo, hi;
    u32 top;
};
GL_HD void acc160_mac(acc160 &acc, u64 a, u64 b) {
    u64 lo = a * b, hi = gl_mulhi64(a, b);
#ifdef __CUDA_ARCH__
    asm("add.cc.u64 %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+l"(acc.lo), "+l"(acc.hi), "+r"(acc.top)
        : "l"(lo), "l"(hi));
#else
    u64 l = acc.lo + lo;
    u64 c = l < lo;
    u64 h = acc.hi + hi;
    u32 c2 = h < hi;
    u64 h2 = h + c;
    c2 += h2 < h;
    acc.lo = l; acc.hi = h2; acc.top += c2;
#endif
}
// top*2^128 + hi*2^64 + lo  (mod p);  2^128 = -2^32
GL_HD u64 acc160_reduce(const acc160 &acc) {
    u64 r = gl_reduce128(acc.lo, acc.hi);
    return gl_sub_c(r, (u64)acc.top << 32);
}

// MDS layer: out[r] = sum_i in[(i+r)%12] * CIRC[i] + in[r]*DIAG[r], on 32-bit halves (no reduction inside):
// 2 x 145 IMAD.WIDE with immediate coefficients (written as PTX so that ptxas keeps them as multiply-adds instead
// of strength-reducing x2/x16 into shift/add chains on the issue-bound alu side), then one 10-instruction fold
// per lane:  al + ah*2^32 = (a0 - h1) + 2^32 * (a1 + h0 + h1)  (mod p),  al = (a1:a0), ah = (h1:h0) < 2^42.
#ifdef __CUDA_ARCH__
// coefficients come from the constant bank (an IMAD.WIDE operand) rather than immediates: with immediates ptxas
// rewrites x16 / x2 as shift pairs and re-associates the chains into 3-input adds (680 instead of 430 instructions)
#define PSD_MACW(acc, x, c) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(c_mds[c]))
#define PSD_MULW(acc, x, c) asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(x), "r"(c_mds[c]))
#define PSD_MDS_ROW(r)                                                                                           \
    {                                                                                                            \
        u64 al, ah;                                                                                              \
        PSD_MULW(al, lo[(0 + r) % 12], 0); PSD_MULW(ah, hi[(0 + r) % 12], 0);                                  \
        PSD_MACW(al, lo[(1 + r) % 12], 1); PSD_MACW(ah, hi[(1 + r) % 12], 1);                                  \
        PSD_MACW(al, lo[(2 + r) % 12], 2); PSD_MACW(ah, hi[(2 + r) % 12], 2);                                  \
        PSD_MACW(al, lo[(3 + r) % 12], 3); PSD_MACW(ah, hi[(3 + r) % 12], 3);                                  \
        PSD_MACW(al, lo[(4 + r) % 12], 4);  PSD_MACW(ah, hi[(4 + r) % 12], 4);                                   \
        PSD_MACW(al, lo[(5 + r) % 12], 5); PSD_MACW(ah, hi[(5 + r) % 12], 5);                                  \
        PSD_MACW(al, lo[(6 + r) % 12], 6); PSD_MACW(ah, hi[(6 + r) % 12], 6);                                  \
        PSD_MACW(al, lo[(7 + r) % 12], 7); PSD_MACW(ah, hi[(7 + r) % 12], 7);                                  \
        PSD_MACW(al, lo[(8 + r) % 12], 8); PSD_MACW(ah, hi[(8 + r) % 12], 8);                                  \
        PSD_MACW(al, lo[(9 + r) % 12], 9); PSD_MACW(ah, hi[(9 + r) % 12], 9);                                  \
        PSD_MACW(al, lo[(10 + r) % 12], 10); PSD_MACW(ah, hi[(10 + r) % 12], 10);                                \
        PSD_MACW(al, lo[(11 + r) % 12], 11); PSD_MACW(ah, hi[(11 + r) % 12], 11);                                \
        if (r == 0) { PSD_MACW(al, lo[0], 12); PSD_MACW(ah, hi[0], 12); }                                          \
        u32 v0, v1;                                                                                              \
        asm("{\n\t"                                                                                              \
            ".reg .u32 a0, a1, h0, h1, t, k, tt, hh;\n\t"                                                        \
            "mov.b64 {a0, a1}, %2;\n\t"                                                                          \
            "mov.b64 {h0, h1}, %3;\n\t"                                                                          \
            "add.u32 t, a1, h1;\n\t"                                                                             \
            "sub.cc.u32 %0, a0, h1;\n\t"                                                                         \
            "subc.cc.u32 %1, h0, 0;\n\t"                                                                         \
            "subc.u32 k, 0, 0;\n\t"                                                                              \
            "add.cc.u32 %1, %1, t;\n\t"                                                                          \
            "addc.u32 k, k, 0;\n\t"                                                                              \
            "sub.u32 tt, 0, k;\n\t"                                                                              \
            "shr.s32 hh, k, 1;\n\t"                                                                              \
            "add.cc.u32 %0, %0, tt;\n\t"                                                                         \
            "addc.u32 %1, %1, hh;\n\t"                                                                           \
            "}"                                                                                                  \
            : "=&r"(v0), "=&r"(v1)                                                                               \
            : "l"(al), "l"(ah));                                                                                 \
        s[r] = ((u64)v1 << 32) | v0;                                                                             \
    }
#endif
// MDS on the FP64 pipe (B200 keeps a full-rate fp64 pipe -- 64 DFMA/clk/SM -- that is otherwise idle in this
// kernel, while the integer alu and fma pipes are the bottleneck; profiles/r01_pipe_model.md).  All products and sums
// of the MDS layer are < 2^42, hence exact in binary64: halves enter as 2^52 + x (bit pattern 0x43300000:x) minus
// 2^52, 2 x 144 DFMA accumulate them, and the sums leave through the mantissa of (sum + 2^52).  Same integers as the
// IMAD.WIDE formulation; fp addition is not re-associated by the compiler, so this is also free of the un-fusing ptxas
// applies to integer multiply-add chains.
#ifndef PSD_MDS_FP64
#define PSD_MDS_FP64 1
#endif
#if defined(__CUDA_ARCH__) && PSD_MDS_FP64
__device__ __forceinline__ void poseidon_mds_fp64(u64 (&s)[12]) {
    const double two52 = 4503599627370496.0;
    const double C[12] = {17., 15., 41., 16., 2., 28., 13., 13., 39., 18., 34., 20.};
    double lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = __hiloint2double(0x43300000, (int)(u32)s[i]) - two52;
        hi[i] = __hiloint2double(0x43300000, (int)(u32)(s[i] >> 32)) - two52;
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double al = lo[r % 12] * C[0], ah = hi[r % 12] * C[0];
#pragma unroll
        for (int i = 1; i < 12; i++) {
            al = fma(lo[(i + r) % 12], C[i], al);
            ah = fma(hi[(i + r) % 12], C[i], ah);
        }
        if (r == 0) {
            al = fma(lo[0], 8., al);
            ah = fma(hi[0], 8., ah);
        }
        const u64 ual = (u64)__double_as_longlong(al + two52) & 0x000FFFFFFFFFFFFFull;
        const u64 uah = (u64)__double_as_longlong(ah + two52) & 0x000FFFFFFFFFFFFFull;
        u32 v0, v1;
        asm("{\n\t"
            ".reg .u32 a0, a1, h0, h1, t, k, tt, hh;\n\t"
            "mov.b64 {a0, a1}, %2;\n\t"
            "mov.b64 {h0, h1}, %3;\n\t"
            "add.u32 t, a1, h1;\n\t"
            "sub.cc.u32 %0, a0, h1;\n\t"
            "subc.cc.u32 %1, h0, 0;\n\t"
            "subc.u32 k, 0, 0;\n\t"
            "add.cc.u32 %1, %1, t;\n\t"
            "addc.u32 k, k, 0;\n\t"
            "sub.u32 tt, 0, k;\n\t"
            "shr.s32 hh, k, 1;\n\t"
            "add.cc.u32 %0, %0, tt;\n\t"
            "addc.u32 %1, %1, hh;\n\t"
            "}"
            : "=&r"(v0), "=&r"(v1)
            : "l"(ual), "l"(uah));
        s[r] = ((u64)v1 << 32) | v0;
    }
}
#endif
GL_HD void poseidon_mds(u64 (&s)[12]) {
#if defined(__CUDA_ARCH__) && PSD_MDS_FP64
    poseidon_mds_fp64(s);
#else
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (u32)s[i];
        hi[i] = (u32)(s[i] >> 32);
    }
#ifdef __CUDA_ARCH__
    PSD_MDS_ROW(0) PSD_MDS_ROW(1) PSD_MDS_ROW(2) PSD_MDS_ROW(3) PSD_MDS_ROW(4) PSD_MDS_ROW(5)
    PSD_MDS_ROW(6) PSD_MDS_ROW(7) PSD_MDS_ROW(8) PSD_MDS_ROW(9) PSD_MDS_ROW(10) PSD_MDS_ROW(11)
#else
    const u32 C[12] = POSEIDON_MDS_CIRC_INIT;
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
        for (int i = 0; i < 12; i++) {
            al += (u64)lo[(i + r) % 12] * C[i];
            ah += (u64)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) {
            al += (u64)lo[0] * POSEIDON_MDS_DIAG0;
            ah += (u64)hi[0] * POSEIDON_MDS_DIAG0;
        }
        // al + ah*2^32, al, ah < 2^42:  ah*2^32 = (ah_lo << 32) + ah_hi*2^64 = (ah_lo << 32) + ah_hi*eps
        u64 l = al + (ah << 32);
        u32 top = (u32)(ah >> 32) + (l < al ? 1u : 0u);
        u64 t = l + (u64)top * GL_EPS;
        s[r] = t + (t < l ? (u64)GL_EPS : 0);
    }
#endif
#endif
}

// Lanes handled per rolled iteration of the full-round S-box loop (must divide 12).  The state is ROTATED by that
// many lanes per iteration so that every iteration addresses the same registers.
#ifndef PSD_SBOX_LANES
#define PSD_SBOX_LANES 3
#endif
GL_HD void poseidon_rot(u64 (&s)[12]) {
    u64 t[PSD_SBOX_LANES];
#pragma unroll
    for (int i = 0; i < PSD_SBOX_LANES; i++) t[i] = s[i];
#pragma unroll
    for (int i = 0; i < 12 - PSD_SBOX_LANES; i++) s[i] = s[i + PSD_SBOX_LANES];
#pragma unroll
    for (int i = 0; i < PSD_SBOX_LANES; i++) s[12 - PSD_SBOX_LANES + i] = t[i];
}

// add round constants + S-box on all 12 lanes
GL_HD void poseidon_full_sbox(u64 (&s)[12], int rc_off) {
#if PSD_SBOX_LANES == 12
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = gl_pow7(gl_add_c(s[k], PSD_RC(rc_off + k)));
#else
    PSD_UNROLL1
    for (int it = 0; it < 12 / PSD_SBOX_LANES; it++) {
#pragma unroll
        for (int k = 0; k < PSD_SBOX_LANES; k++) s[k] = gl_pow7(gl_add_c(s[k], PSD_RC(rc_off + PSD_SBOX_LANES * it + k)));
        poseidon_rot(s);
    }
#endif
}

GL_HD void poseidon_partial_rounds(u64 (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], PSD_FIRST(i));
    {   // dense 11x11 matrix on lanes 1..11 (lane 0 unchanged); rows rolled, outputs shifted through o[]
        u64 o[11];
#pragma unroll
        for (int i = 0; i < 11; i++) o[i] = 0;
        PSD_UNROLL1
        for (int i = 0; i < 11; i++) {
            acc160 acc = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < 11; j++) acc160_mac(acc, PSD_INIT(11 * i + j), s[j + 1]);
            u64 v = acc160_reduce(acc);
#pragma unroll
            for (int k = 0; k < 10; k++) o[k] = o[k + 1];
            o[10] = v;
        }
#pragma unroll
        for (int i = 0; i < 11; i++) s[i + 1] = o[i];
    }
    PSD_UNROLL1
    for (int r = 0; r < 22; r++) {
        u64 s0 = gl_add_c(gl_pow7(s[0]), PSD_K(r));
        acc160 acc = {0, 0, 0};
        acc160_mac(acc, s0, 25);  // MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0]
#pragma unroll
        for (int i = 0; i < 11; i++) acc160_mac(acc, PSD_ROW(11 * r + i), s[i + 1]);
#pragma unroll
        for (int i = 0; i < 11; i++) s[i + 1] = gl_mul_add(PSD_COL(11 * r + i), s0, s[i + 1]);
        s[0] = acc160_reduce(acc);
    }
}

// The permutation on the integer pipes only (host replay, and the A/B baseline of the FP64 formulation).
GL_HD void poseidon_permute_int(u64 (&s)[12]) {
    PSD_UNROLL1
    for (int r = 0; r < 8; r++) {
        if (r == 4) poseidon_partial_rounds(s);
        poseidon_full_sbox(s, 12 * (r < 4 ? r : r + 22));
        poseidon_mds(s);
    }
}

#ifndef PSD_F64
#define PSD_F64 1
#endif
// The permutation.  Accepts non-canonical lanes; outputs are exact residues, not necessarily canonical.
// Device: linear layers on the FP64 pipe (poseidon_f64.cuh); the integer-only form is kept for the CPU replay harness.
GL_HD void poseidon_permute(u64 (&s)[12]) {
#if defined(__CUDA_ARCH__) && PSD_F64
    poseidon_permute_f64(s);
#else
    poseidon_permute_int(s);
#endif
}

// two_to_one(l, r) = permute([l, r, 0, 0, 0, 0])[0..4]
GL_HD void poseidon_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_canon(s[i]);
}
