// Poseidon permutation with every LINEAR layer on the FP64 pipe (exact integer arithmetic below 2^53).
//
// Same function as plonky2::hash::poseidon::Poseidon::poseidon for GoldilocksField (dep plonky2 0.1.4,
// /root/reference/Cargo.lock:2347-2350): 4 full + 22 partial + 4 full rounds of
//     state += RC_t;  S-box (x^7) on all lanes / on lane 0;  state = MDS * state.
//
// Why (profiles/r01_pipe_model.md, profiles/r01_leaves_v1.md): on sm_100a the integer alu pipe is half rate and is
// what bounds a Poseidon kernel written on the integer pipes (the "fast" partial rounds of plonky2 are 23 full
// 64x64 products with 64-bit constants per round: 852 alu/fma issue cycles per round per warp).  B200 keeps a
// 64-lane/clk/SM FP64 pipe that such a kernel leaves idle.  Here
//   * a lane is carried as a pair of doubles (lo, hi) meaning lo + 2^32*hi (mod p), both signed integers < 2^51;
//   * the S-box runs on the integer pipes (3 reduced products + 1 unreduced 128-bit product x0..x3); the final
//     reduction 2^64 = 2^32 - 1, 2^96 = -1 is left to the pair:  lo = x0 - x2 - x3,  hi = x1 + x2  (two small integer
//     sums, one exact conversion each);
//   * the MDS layer (coefficients <= 41) is a 12-point circular correlation computed as a split convolution (pf_circ12:
//     90 FP64 operations per half instead of 144); the NEXT round's constants are the initial values of the accumulation
//     chains, so adding round constants costs nothing;
//   * the 22 partial rounds run in their NAIVE form, two rounds per step: with W = state after the first S-box,
//         b  = (M W + RC_{t+1})[0],   b' = b^7,   d = b' - b,
//         state_{t+2} = M^2 W + (M RC_{t+1} + RC_{t+2}) + d * M e_0
//                     = C^2 W + K + (8 W_0 + d) * C e_0 + 8 b' e_0          (M = C + 8 e_0 e_0^T, C = circ)
//     (C^2 has entries < 2^15, so 2^32 * 264^2 < 2^49 stays exact).  Lanes 1..11 never leave the FP64 domain during
//     the partial rounds; they are re-normalised to |lo|,|hi| <= 2^31 + 2^19 with 8 FP64 operations every two rounds.
//     Only lane 0 crosses to the integer side (one fold per round) for its S-box;
//   * a pair (al, ah) is folded back to a u64 through the mantissa of al + 1.5*2^52 (a bias of 2^51 on both halves,
//     compensated inside the chain-init constants), then one 10-instruction carry chain.
//
// Everything is exact: the result is bit-identical to the integer formulation (tests: upstream KATs, oracle parity).
// The file is __host__ __device__ so that tests/emu replays the same arithmetic on the CPU of the GPU-less
// development container (IEEE binary64 with fma is the same arithmetic on both sides).
#pragma once
#include "gl64.cuh"
#include "poseidon_consts.h"
#include <math.h>
#include <string.h>

struct PsdF64Tables {
    double full_init[8][12][2];  // chain-init constants of the 8 full-round MDS layers [layer][lane][lo, hi]
    double pair_t0[11][2];       // RC_{t+1}[0] - B
    double pair_k[11][12][2];    // M RC_{t+1} + RC_{t+2} - B * M e_0 (- B where the lane is folded next)
    double m_row0[12];           // M[0][j]
    double circ[12];             // MDS_MATRIX_CIRC
    // split-convolution form (PF_SPLIT): a 12-point circular correlation as (6 cyclic + 6 negacyclic), the cyclic half
    // again as (3 + 3): 90 FP64 operations instead of 144.  [0..3) = cPP, [3..6) = cPQ, [6..12) = cQ
    double sc1[12];              // for circ(CIRC)
    double sc2[12];              // for circ(CIRC)^2
    double c_col0[12];           // CIRC[(12 - r) % 12]:  C e_0
    double full_init_s[8][2][12];  // chain-init constants in split form [layer][half][UU(3), UV(3), V(6)]
    double pair_k_s[11][2][12];    // same for the pair constants (lane 0 also carries -8 * pair_t0)
};

#ifdef __CUDACC__
__constant__ PsdF64Tables c_pf;
#endif
#ifdef __CUDA_ARCH__
#define PF_T(x) c_pf.x
#else
static PsdF64Tables h_pf;
#define PF_T(x) h_pf.x
#endif

#define PF_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */
#define PF_BIAS 0x000FFFFFFFF80000ULL /* B = 2^51 + 2^83 = 2^52 - 2^19 (mod p): what a fold adds to the value */

// ---- host: table construction (exact field arithmetic) ----
static inline void psd_f64_build_tables(PsdF64Tables &t) {
    const u64 circ[12] = POSEIDON_MDS_CIRC_INIT;
    u64 M[12][12];
    memset(M, 0, sizeof(M));
    for (int r = 0; r < 12; r++)
        for (int i = 0; i < 12; i++) M[r][(i + r) % 12] += circ[i];
    M[0][0] += POSEIDON_MDS_DIAG0;
    for (int j = 0; j < 12; j++) {
        t.m_row0[j] = (double)M[0][j];
        t.circ[j] = (double)circ[j];
    }
    auto sub = [](u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); };   // canonical a, b
    auto add = [](u64 a, u64 b) { u64 s = a + b; return (s < a || s >= GL_P) ? s - GL_P : s; };
    auto put = [](double (&d)[2], u64 v) { d[0] = (double)(u32)v; d[1] = (double)(u32)(v >> 32); };
    auto rc = [](int round, int lane) -> u64 { return round < 30 ? POSEIDON_RC[12 * round + lane] % GL_P : 0; };
    // split-convolution constants of a circulant first row c:  cP = (c[i] + c[i+6]) / 2, cQ = (c[i] - c[i+6]) / 2,
    // cPP = (cP[i] + cP[i+3]) / 2, cPQ = (cP[i] - cP[i+3]) / 2   (halves and quarters are exact in binary64)
    auto split_row = [](const double (&c)[12], double (&out)[12]) {
        double cP[6];
        for (int i = 0; i < 6; i++) { cP[i] = (c[i] + c[i + 6]) / 2; out[6 + i] = (c[i] - c[i + 6]) / 2; }
        for (int i = 0; i < 3; i++) { out[i] = (cP[i] + cP[i + 3]) / 2; out[3 + i] = (cP[i] - cP[i + 3]) / 2; }
    };
    {
        double c1[12], c2[12];
        for (int m = 0; m < 12; m++) {
            c1[m] = (double)circ[m];
            u64 a = 0;
            for (int i = 0; i < 12; i++) a += circ[i] * circ[(m - i + 12) % 12];
            c2[m] = (double)a;
            t.c_col0[m] = (double)circ[(12 - m) % 12];
        }
        split_row(c1, t.sc1);
        split_row(c2, t.sc2);
    }
    const u64 B = PF_BIAS;
    for (int L = 0; L < 8; L++) {
        int next_round = L < 4 ? L + 1 : 27 + (L - 4);   // constants of the round that follows the layer
        for (int j = 0; j < 12; j++) put(t.full_init[L][j], sub(rc(next_round, j), B));
    }
    for (int p = 0; p < 11; p++) {
        int r1 = 4 + 2 * p + 1, r2 = r1 + 1;
        put(t.pair_t0[p], sub(rc(r1, 0), B));
        for (int r = 0; r < 12; r++) {
            u64 k = rc(r2, r);
            for (int j = 0; j < 12; j++) k = add(k, h_gl_mul(M[r][j], rc(r1, j)));
            k = sub(k, h_gl_mul(B, M[r][0]));
            if (r == 0 || p == 10) k = sub(k, B);
            put(t.pair_k[p][r], k);
        }
    }
    // the same chain-init constants in split form (same transform as the coefficients)
    for (int L = 0; L < 8; L++)
        for (int h = 0; h < 2; h++) {
            double k[12];
            for (int j = 0; j < 12; j++) k[j] = t.full_init[L][j][h];
            split_row(k, t.full_init_s[L][h]);
        }
    for (int p = 0; p < 11; p++)
        for (int h = 0; h < 2; h++) {
            double k[12];
            for (int j = 0; j < 12; j++) k[j] = t.pair_k[p][j][h];
            k[0] -= 8.0 * t.pair_t0[p][h];   // lane 0 receives 8 * T0, T0 = (C W)[0] + 8 W_0 + pair_t0
            split_row(k, t.pair_k_s[p][h]);
        }
}

#ifdef __CUDACC__
static inline cudaError_t psd_f64_upload_tables() {
    PsdF64Tables t;
    psd_f64_build_tables(t);
    return cudaMemcpyToSymbol(c_pf, &t, sizeof(t));
}
#endif

// ---- primitives ----
// PF_TRACK_FOLD / PF_TRACK_RENORM: host-only hooks of the CPU replay (tests/emu) recording the range of the values that
// enter a fold (needs -2^51 <= x < 2^51) and a re-normalisation (accumulated sums, < 2^50).  Nothing in the product.
#ifndef PF_TRACK_FOLD
#define PF_TRACK_FOLD(x)
#define PF_TRACK_RENORM(x)
#endif
#ifndef PF_INT_CVT
#define PF_INT_CVT 1     // bit 0 = S-box output (faster, default), bit 1 = fold (slower) through 64-bit integer conversions instead of DADDs
#endif
#ifndef PF_CVT_MAGIC
#define PF_CVT_MAGIC 0   // 0: I2F.F64.U32 on the conversion unit; 1: 2^52-mantissa trick (one more DADD, two moves)
#endif
GL_HD double pf_cvt(u32 w) {
#if defined(__CUDA_ARCH__) && PF_CVT_MAGIC
    return __hiloint2double(0x43300000, (int)w) - 4503599627370496.0;
#else
    return (double)w;
#endif
}
GL_HD double pf_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return fma(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
GL_HD u64 pf_bits(double x) {
#ifdef __CUDA_ARCH__
    return (u64)__double_as_longlong(x);
#else
    u64 u;
    memcpy(&u, &x, 8);
    return u;
#endif
}

// (al, ah), both in (-2^51, 2^51)  ->  some u64 = (al + 2^51) + 2^32 * (ah + 2^51)  (mod p).
// With ua = a0 + 2^32 a1, uh = h0 + 2^32 h1 (a1, h1 < 2^20):  value = (a0 - h1) + 2^32 * (a1 + h0 + h1).
GL_HD u64 pf_fold(double al, double ah) {
    PF_TRACK_FOLD(al);
    PF_TRACK_FOLD(ah);
#if defined(__CUDA_ARCH__) && (PF_INT_CVT & 2)
    const u64 ua = (u64)(__double2ll_rz(al) + (1ll << 51)), uh = (u64)(__double2ll_rz(ah) + (1ll << 51));
#else
    const u64 ua = pf_bits(al + PF_MAGIC), uh = pf_bits(ah + PF_MAGIC);   // mantissa = x + 2^51
#endif
#ifdef __CUDA_ARCH__
    u32 v0, v1;
    asm("{\n\t"
        ".reg .u32 a0, a1, h0, h1, t, k, tt, hh;\n\t"
        "mov.b64 {a0, a1}, %2;\n\t"
        "mov.b64 {h0, h1}, %3;\n\t"
        "and.b32 a1, a1, 0xFFFFF;\n\t"
        "and.b32 h1, h1, 0xFFFFF;\n\t"
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
        : "l"(ua), "l"(uh));
    return ((u64)v1 << 32) | v0;
#else
    const u64 a = ua & 0xFFFFFFFFFFFFFULL, h = uh & 0xFFFFFFFFFFFFFULL;
    u64 l = a + (h << 32);
    u64 top = (h >> 32) + (l < a ? 1u : 0u);
    u64 t = l + top * GL_EPS;
    return t + (t < l ? (u64)GL_EPS : 0);
#endif
}

// re-normalise a pair without changing lo + 2^32 hi (mod p):  |lo|, |hi| < 2^50  ->  |lo| <= 2^31 + 2^18, |hi| <= 2^31 + 2^19
GL_HD void pf_renorm(double &lo, double &hi) {
    PF_TRACK_RENORM(lo);
    PF_TRACK_RENORM(hi);
    const double I32 = 1.0 / 4294967296.0, N32 = -4294967296.0;
    const double a1 = pf_fma(lo, I32, PF_MAGIC) - PF_MAGIC;   // rint(lo / 2^32)
    const double a0 = pf_fma(a1, N32, lo);
    const double h1 = pf_fma(hi, I32, PF_MAGIC) - PF_MAGIC;
    const double h01 = pf_fma(h1, -4294967295.0, hi);          // h0 + h1 in one operation (h0 = hi - 2^32 h1)
    lo = a0 - h1;                                              // 2^64 h1 = 2^32 h1 - h1
    hi = a1 + h01;
}

// x^7 with the last product left unreduced (x0 + 2^32 x1 + 2^64 x2 + 2^96 x3) and reduced in FP64.
GL_HD void pf_pow7(u64 x, double &lo, double &hi) {
    const u64 x2 = gl_sqr(x);
    const u64 x4 = gl_sqr(x2);
    const u64 x3 = gl_mul(x, x2);
    const u64 pl = x3 * x4, ph = gl_mulhi64(x3, x4);
#if (PF_INT_CVT & 1)
    // the two limb sums as 64-bit integers (|lo| < 2^34, hi < 2^33), one exact conversion each: 2 I2F.F64.S64 + 4 integer
    // instructions instead of 4 I2F.F64.U32 + 3 DADD (0.9 % faster on sm_100a: the FP64 pipe is the busier one)
    lo = (double)((long long)(u64)(u32)pl - (long long)(u64)(u32)ph - (long long)(ph >> 32));
    hi = (double)(long long)((pl >> 32) + (u64)(u32)ph);
#else
    const double d0 = pf_cvt((u32)pl), d1 = pf_cvt((u32)(pl >> 32)), d2 = pf_cvt((u32)ph), d3 = pf_cvt((u32)(ph >> 32));
    lo = (d0 - d2) - d3;
    hi = d1 + d2;
#endif
}

// y = init + circ(c) x as a split convolution (see PsdF64Tables::sc1).  cc = split coefficients, init = split constants.
GL_HD void pf_circ12(const double (&x)[12], const double *cc, const double *init, double (&y)[12]) {
    double P[6], Q[6], PP[3], PQ[3], U[6];
#pragma unroll
    for (int j = 0; j < 6; j++) { P[j] = x[j] + x[j + 6]; Q[j] = x[j] - x[j + 6]; }
#pragma unroll
    for (int j = 0; j < 3; j++) { PP[j] = P[j] + P[j + 3]; PQ[j] = P[j] - P[j + 3]; }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double uu = init[r], uv = init[3 + r];
#pragma unroll
        for (int i = 0; i < 3; i++) uu = pf_fma(PP[(i + r) % 3], cc[i], uu);
#pragma unroll
        for (int i = 0; i < 3; i++) uv = (i + r >= 3) ? pf_fma(-PQ[(i + r) % 3], cc[3 + i], uv) : pf_fma(PQ[(i + r) % 3], cc[3 + i], uv);
        U[r] = uu + uv;
        U[r + 3] = uu - uv;
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double v = init[6 + r];
#pragma unroll
        for (int i = 0; i < 6; i++) v = (i + r >= 6) ? pf_fma(-Q[(i + r) % 6], cc[6 + i], v) : pf_fma(Q[(i + r) % 6], cc[6 + i], v);
        y[r] = U[r] + v;
        y[r + 6] = U[r] - v;
    }
}

#ifndef PF_SBOX_LANES
#define PF_SBOX_LANES 12
#endif
#ifdef __CUDA_ARCH__
#define PF_UNROLL1 _Pragma("unroll 1")
#else
#define PF_UNROLL1
#endif

// One full round on an integer state that already holds state + RC_t:  S-box, MDS (+ RC_{t+1} - B), fold.
GL_HD void pf_full_round(u64 (&s)[12], int layer) {
    double xl[12], xh[12];
#if PF_SBOX_LANES == 12
#pragma unroll
    for (int k = 0; k < 12; k++) pf_pow7(s[k], xl[k], xh[k]);
#else
    // rolled: PF_SBOX_LANES lanes per iteration; s, xl, xh rotate so that every iteration addresses the same registers
#pragma unroll
    for (int k = 0; k < 12; k++) xl[k] = xh[k] = 0.0;
    PF_UNROLL1
    for (int it = 0; it < 12 / PF_SBOX_LANES; it++) {
        double tl[PF_SBOX_LANES], th[PF_SBOX_LANES];
#pragma unroll
        for (int k = 0; k < PF_SBOX_LANES; k++) pf_pow7(s[k], tl[k], th[k]);
#pragma unroll
        for (int k = 0; k < 12 - PF_SBOX_LANES; k++) {
            s[k] = s[k + PF_SBOX_LANES];
            xl[k] = xl[k + PF_SBOX_LANES];
            xh[k] = xh[k + PF_SBOX_LANES];
        }
#pragma unroll
        for (int k = 0; k < PF_SBOX_LANES; k++) {
            xl[12 - PF_SBOX_LANES + k] = tl[k];
            xh[12 - PF_SBOX_LANES + k] = th[k];
        }
    }
#endif
    double al[12], ah[12];
    pf_circ12(xl, PF_T(sc1), PF_T(full_init_s)[layer][0], al);
    pf_circ12(xh, PF_T(sc1), PF_T(full_init_s)[layer][1], ah);
    al[0] = pf_fma(xl[0], 8.0, al[0]);
    ah[0] = pf_fma(xh[0], 8.0, ah[0]);
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = pf_fold(al[r], ah[r]);
}

// The 22 partial rounds, two per step.  In: integer state + RC_4.  Out: integer state + RC_26.
GL_HD void pf_partial_rounds(u64 (&s)[12]) {
    double al[12], ah[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        al[j] = pf_cvt((u32)s[j]);
        ah[j] = pf_cvt((u32)(s[j] >> 32));
    }
    al[0] -= 2251799813685248.0;   // lane 0 enters every step through a fold (which adds 2^51 to both halves)
    ah[0] -= 2251799813685248.0;
    PF_UNROLL1
    for (int p = 0; p < 11; p++) {
        const u64 a = pf_fold(al[0], ah[0]);
        if (p != 0) {                                   // step 0 starts from fresh 32-bit limbs
#pragma unroll
            for (int j = 1; j < 12; j++) pf_renorm(al[j], ah[j]);
        }
        pf_pow7(a, al[0], ah[0]);                       // W
        double t0l = PF_T(pair_t0)[p][0], t0h = PF_T(pair_t0)[p][1];
        // T0 = (C W)[0] + 8 W_0 + const;  M^2 W = circ(c2) W + W_0 * (8 C e_0) + e_0 * 8 (T0 - const)
#pragma unroll
        for (int j = 0; j < 12; j++) {                  // row 0 of M = circ + 8 e_0 e_0^T
            t0l = pf_fma(al[j], PF_T(m_row0)[j], t0l);
            t0h = pf_fma(ah[j], PF_T(m_row0)[j], t0h);
        }
        const u64 b = pf_fold(t0l, t0h);                // lane 0 entering the second S-box (RC included)
        double nl[12], nh[12];                          // M^2 W + K: independent of b, overlaps the S-box below
        pf_circ12(al, PF_T(sc2), PF_T(pair_k_s)[p][0], nl);
        pf_circ12(ah, PF_T(sc2), PF_T(pair_k_s)[p][1], nh);
        double bl, bh;
        pf_pow7(b, bl, bh);
        // the two rank-1 terms share their direction: W_0 * 8 C e_0 + d * (C e_0 + 8 e_0) = (8 W_0 + d) * C e_0 + 8 d e_0
        // with d = (b' - b) + B as a pair, and on lane 0  8 T0 + 8 d = 8 b'
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
    for (int r = 0; r < 12; r++) s[r] = pf_fold(al[r], ah[r]);
}

// The permutation.  Accepts non-canonical lanes; outputs are exact residues, not necessarily canonical.
GL_HD void poseidon_permute_f64(u64 (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], POSEIDON_RC_AT(i));
    PF_UNROLL1
    for (int L = 0; L < 8; L++) {
        if (L == 4) pf_partial_rounds(s);
        pf_full_round(s, L);
    }
}
