// Poseidon over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds), one sponge state per thread.
//
// Replaces plonky2::hash::poseidon::Poseidon::poseidon for GoldilocksField and the sponge helpers of
// plonky2::hash::hashing (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350).
// Same outputs as the reference ("naive") round function; the partial rounds run in the algebraically
// identical sparse form (constants derived by tools/gen_poseidon_consts.py and checked against the
// upstream known-answer vectors).
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
#endif
#ifdef __CUDA_ARCH__
#define PSD_RC(i) c_rc[i]
#define PSD_FIRST(i) c_fast_first[i]
#define PSD_K(i) c_fast_k[i]
#define PSD_ROW(i) c_fast_row[i]
#define PSD_COL(i) c_fast_col[i]
#define PSD_INIT(i) c_fast_init[i]
#else
#define PSD_RC(i) POSEIDON_RC[i]
#define PSD_FIRST(i) POSEIDON_FAST_FIRST[i]
#define PSD_K(i) POSEIDON_FAST_K[i]
#define PSD_ROW(i) POSEIDON_FAST_ROW[i]
#define PSD_COL(i) POSEIDON_FAST_COL[i]
#define PSD_INIT(i) POSEIDON_FAST_INIT[i]
#endif

#ifdef __CUDACC__
// Uploads the constant tables of this translation unit; call once per module before the first launch.
static inline cudaError_t poseidon_upload_constants() {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_rc, POSEIDON_RC, sizeof(POSEIDON_RC))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_first, POSEIDON_FAST_FIRST, sizeof(POSEIDON_FAST_FIRST))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_k, POSEIDON_FAST_K, sizeof(POSEIDON_FAST_K))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_row, POSEIDON_FAST_ROW, sizeof(POSEIDON_FAST_ROW))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_fast_col, POSEIDON_FAST_COL, sizeof(POSEIDON_FAST_COL))) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_fast_init, POSEIDON_FAST_INIT, sizeof(POSEIDON_FAST_INIT));
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

// MDS layer: out[r] = sum_i in[(i+r)%12] * CIRC[i] + in[r]*DIAG[r], on 32-bit halves (no reduction inside).
GL_HD void poseidon_mds(u64 (&s)[12]) {
    const u32 C[12] = POSEIDON_MDS_CIRC_INIT;
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (u32)s[i];
        hi[i] = (u32)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
#pragma unroll
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
}

GL_HD void poseidon_full_round(u64 (&s)[12], int rc_off) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add_c(s[i], PSD_RC(rc_off + i)));
    poseidon_mds(s);
}

GL_HD void poseidon_partial_rounds(u64 (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], PSD_FIRST(i));
    {   // dense 11x11 matrix on lanes 1..11 (lane 0 unchanged)
        u64 o[11];
#pragma unroll
        for (int i = 0; i < 11; i++) {
            acc160 acc = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < 11; j++) acc160_mac(acc, PSD_INIT(11 * i + j), s[j + 1]);
            o[i] = acc160_reduce(acc);
        }
#pragma unroll
        for (int i = 0; i < 11; i++) s[i + 1] = o[i];
    }
#pragma unroll 1
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

// The permutation.  Accepts non-canonical lanes; outputs are exact residues, not necessarily canonical.
GL_HD void poseidon_permute(u64 (&s)[12]) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) poseidon_full_round(s, 12 * r);
    poseidon_partial_rounds(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) poseidon_full_round(s, 12 * (4 + 22 + r));
}

// two_to_one(l, r) = permute([l, r, 0, 0, 0, 0])[0..4]
GL_HD void poseidon_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl_canon(s[i]);
}
