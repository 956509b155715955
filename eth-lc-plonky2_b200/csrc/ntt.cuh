// Goldilocks NTT passes: iNTT (values -> coefficients) and coset LDE (coefficients -> 2^r cosets).
//
// Replaces the per-column fft_classic of plonky2_field::fft and the "IFFT" / "FFT + blinding" / "transpose
// LDEs" stages of plonky2::fri::oracle::PolynomialBatch::{from_values, from_coeffs} (dep plonky2 0.1.4,
// /root/reference/Cargo.lock:2347-2350, reached from /root/reference/eth-lc-plonky2/src/main.rs:227,230;
// SURVEY.md 3.3, A.2, A.3).
//
// B200 design (not plonky2's): a size-n transform is at most two passes over HBM.  A pass stages a tile of
// P = 2^LOGP points x A adjacent lanes in shared memory (every global access is an 8*A-byte segment or a
// contiguous run), runs the P-point decimation-in-frequency network with radix-16 register blocks, applies the
// four-step twiddle on the way out and writes each element exactly once.
//
//   LDE:  leaves are kept COLUMN-major in bit-reversed row order, lde[c][k] = f_c(7 * w_L^{bitrev(k)}).  In that
//         order the 2^r cosets are contiguous blocks: block b (rows [b*n, (b+1)*n)) is the size-n DIF transform
//         (natural in, bit-reversed out) of c_i * (7 * w_L^{bitrev_r(b)})^i, so plonky2's transpose +
//         reverse_index_bits pass disappears and the Merkle leaf hash reads coalesced columns.
//   iNTT: natural in, natural out by the four-step index map (pass 1 writes Y[n1][k2], pass 2 reads it strided).
#pragma once
#include "gl64.cuh"

#define NTT_MAX_LOGP 12

enum NttMode : int {
    NTT_LDE_FIRST = 0,   // coeffs (strided tile) * shift powers -> DIF -> * w_n^{r k1} -> block b, in-place order
    NTT_LDE_SINGLE = 1,  // n <= 2^LOGP: whole coset transform in one pass
    NTT_DIF_LAST = 2,    // contiguous P-point DIF blocks, in place, no twiddle
    NTT_INTT_P1 = 3,     // values (strided tile) -> inverse DIF -> * w_n^{-n1 k2} -> Y[n1][k2]
    NTT_INTT_P2 = 4,     // Y (strided tile) -> inverse DIF -> * 1/n -> X[N2 k1 + k2]
    NTT_INTT_SINGLE = 5, // n <= 2^LOGP
};

struct NttPass {
    const u64 *in;
    u64 *out;
    u32 log_n;         // size of one transform (2^log_n points)
    u32 log_p;         // points handled by this pass
    u32 log_a;         // lanes per tile
    u32 rate_bits;     // LDE only
    u32 num_cols;      // columns (polynomials) in the batch
    u64 in_col_stride;   // elements between columns of `in`
    u64 out_col_stride;  // elements between columns of `out`
    u64 num_tiles;     // total work items
    const u64 *tw_local;  // w_P^e (or inverse), e < P/2
    const u64 *w_lo, *w_hi;  // w_n^e = w_hi[e >> w_lo_bits] * w_lo[e & mask]   (or inverse powers)
    u32 w_lo_bits;
    const u64 *shift_a, *shift_b;  // LDE: s_e^{j*st} [e][P]  and  s_e^{r} [e][st];  s_e = 7 * w_L^e
    u64 scale;         // iNTT: 1/n
};

GL_HD u32 ntt_brev(u32 x, u32 bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

GL_HD u32 ntt_pitch(u32 log_p) { u32 P = 1u << log_p; return P + (P >> 4) + 1; }
GL_HD u32 ntt_sm(u32 pitch, u32 a, u32 j) { return a * pitch + j + (j >> 4); }
static inline size_t ntt_smem_bytes(u32 log_p, u32 log_a) { return (size_t)ntt_pitch(log_p) * (1u << log_a) * sizeof(u64); }

GL_HD u64 ntt_twiddle2(const NttPass &p, u64 e) {
    u64 lo = p.w_lo[e & ((1ull << p.w_lo_bits) - 1)];
    u64 hi = p.w_hi[e >> p.w_lo_bits];
    return gl_mul(lo, hi);
}

// ---- phase 1: global -> shared ----
template <int MODE>
GL_HD void ntt_load(const NttPass &p, u64 *sm, u64 tile, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, A = 1u << p.log_a, pitch = ntt_pitch(p.log_p);
    const u32 total = P << p.log_a;
    if (MODE == NTT_LDE_FIRST || MODE == NTT_INTT_P1 || MODE == NTT_INTT_P2) {
        // strided tile: element (j, a) = in[col][j*st + r0 + a]
        const u32 log_st = p.log_n - p.log_p;
        const u64 tiles_per_col = (MODE == NTT_LDE_FIRST) ? ((u64)(1u << log_st) >> p.log_a) << p.rate_bits
                                                           : ((u64)(1u << log_st) >> p.log_a);
        const u64 col = tile / tiles_per_col;
        u64 rem = tile % tiles_per_col;
        u32 e = 0;
        if (MODE == NTT_LDE_FIRST) { e = (u32)(rem & ((1u << p.rate_bits) - 1)); rem >>= p.rate_bits; }
        const u64 r0 = rem << p.log_a;
        const u64 *src = p.in + col * p.in_col_stride;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 a = idx & (A - 1), j = idx >> p.log_a;
            u64 v = src[((u64)j << log_st) + r0 + a];
            if (MODE == NTT_LDE_FIRST) {
                u64 s = gl_mul(p.shift_a[((u64)e << p.log_p) + j], p.shift_b[((u64)e << log_st) + r0 + a]);
                v = gl_mul(v, s);
            }
            sm[ntt_sm(pitch, a, j)] = v;
        }
    } else if (MODE == NTT_LDE_SINGLE) {
        // lane = (col, e); element j = coeffs[col][j] * s_e^j
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 j = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;  // col * 2^r + e
            u64 col = unit >> p.rate_bits;
            u32 e = (u32)(unit & ((1u << p.rate_bits) - 1));
            u64 v = 0;
            if (col < p.num_cols) v = gl_mul(p.in[col * p.in_col_stride + j], p.shift_b[((u64)e << p.log_p) + j]);
            sm[ntt_sm(pitch, a, j)] = v;
        }
    } else {
        // NTT_DIF_LAST / NTT_INTT_SINGLE: lane a = a-th consecutive run of P contiguous elements
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 j = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            u64 v = 0;
            if (MODE == NTT_DIF_LAST) {
                // units tile the LDE buffer column by column: unit = col * (L/P) + block-in-column
                u32 lpc = p.log_n + p.rate_bits - p.log_p;
                u64 col = unit >> lpc, off = (unit & (((u64)1 << lpc) - 1)) << p.log_p;
                if (col < p.num_cols) v = p.in[col * p.in_col_stride + off + j];
            } else {
                if (unit < p.num_cols) v = p.in[unit * p.in_col_stride + j];
            }
            sm[ntt_sm(pitch, a, j)] = v;
        }
    }
}

// ---- phase 2: one register-blocked round of R DIF stages (stages t0+1 .. t0+R of the P-point network) ----
template <int R>
GL_HD void ntt_round(const NttPass &p, u64 *sm, u32 t0, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, pitch = ntt_pitch(p.log_p);
    const u32 blocks_per_lane = P >> R;
    const u32 nblocks = blocks_per_lane << p.log_a;
    const u32 log_js = p.log_p - t0 - R;  // log2 of the element stride inside the register block
    for (u32 blk = tid; blk < nblocks; blk += nthreads) {
        u32 a = blk / blocks_per_lane, T = blk % blocks_per_lane;
        u32 lo = T & ((1u << log_js) - 1), hi = T >> log_js;
        u32 jb = (hi << (p.log_p - t0)) + lo;
        u64 v[1 << R];
#pragma unroll
        for (int m = 0; m < (1 << R); m++) v[m] = sm[ntt_sm(pitch, a, jb + ((u32)m << log_js))];
#pragma unroll
        for (int u = 1; u <= R; u++) {
            const int span = 1 << (R - u);
#pragma unroll
            for (int m = 0; m < (1 << R); m++) {
                if (m & span) continue;
                // pair (j, j + half), half = span << log_js; twiddle w_P^{(j mod half) << (t0+u-1)}
                u32 jm = (((u32)m & (span - 1)) << log_js) + lo;
                u64 w = p.tw_local[jm << (t0 + u - 1)];
                u64 x = v[m], y = v[m + span];
                v[m] = gl_add(x, y);
                v[m + span] = gl_mul(gl_sub(x, y), w);
            }
        }
#pragma unroll
        for (int m = 0; m < (1 << R); m++) sm[ntt_sm(pitch, a, jb + ((u32)m << log_js))] = v[m];
    }
}

// ---- phase 3: shared -> global; slot q of a lane holds frequency brev(q) ----
template <int MODE>
GL_HD void ntt_store(const NttPass &p, const u64 *sm, u64 tile, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, A = 1u << p.log_a, pitch = ntt_pitch(p.log_p);
    const u32 total = P << p.log_a;
    const u32 log_st = p.log_n - p.log_p;
    if (MODE == NTT_LDE_FIRST) {
        const u64 tiles_per_col = ((u64)(1u << log_st) >> p.log_a) << p.rate_bits;
        const u64 col = tile / tiles_per_col;
        u64 rem = tile % tiles_per_col;
        const u32 e = (u32)(rem & ((1u << p.rate_bits) - 1));
        const u64 r0 = (rem >> p.rate_bits) << p.log_a;
        const u32 b = ntt_brev(e, p.rate_bits);
        u64 *dst = p.out + col * p.out_col_stride + ((u64)b << p.log_n);
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 a = idx & (A - 1), q = idx >> p.log_a;
            u64 w = ntt_twiddle2(p, (r0 + a) * (u64)ntt_brev(q, p.log_p));
            dst[((u64)q << log_st) + r0 + a] = gl_mul(sm[ntt_sm(pitch, a, q)], w);
        }
    } else if (MODE == NTT_LDE_SINGLE) {
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 q = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            u64 col = unit >> p.rate_bits;
            u32 e = (u32)(unit & ((1u << p.rate_bits) - 1));
            u32 b = ntt_brev(e, p.rate_bits);
            if (col < p.num_cols) p.out[col * p.out_col_stride + ((u64)b << p.log_n) + q] = gl_canon(sm[ntt_sm(pitch, a, q)]);
        }
    } else if (MODE == NTT_DIF_LAST) {
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 q = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            u32 lpc = p.log_n + p.rate_bits - p.log_p;
            u64 col = unit >> lpc, off = (unit & (((u64)1 << lpc) - 1)) << p.log_p;
            if (col < p.num_cols) p.out[col * p.out_col_stride + off + q] = gl_canon(sm[ntt_sm(pitch, a, q)]);
        }
    } else if (MODE == NTT_INTT_P1) {
        // Y[col][(r0+a) * P + k2], k2 = brev(q), times w_n^{-(r0+a) k2}
        const u64 tiles_per_col = (u64)(1u << log_st) >> p.log_a;
        const u64 col = tile / tiles_per_col;
        const u64 r0 = (tile % tiles_per_col) << p.log_a;
        u64 *dst = p.out + col * p.out_col_stride;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 k2 = idx & (P - 1), a = idx >> p.log_p;
            u64 w = ntt_twiddle2(p, (r0 + a) * (u64)k2);
            dst[((r0 + a) << p.log_p) + k2] = gl_mul(sm[ntt_sm(pitch, a, ntt_brev(k2, p.log_p))], w);
        }
    } else if (MODE == NTT_INTT_P2) {
        // X[col][k1 * N2 + k0 + a], k1 = brev(q), N2 = 2^log_st, times 1/n
        const u64 tiles_per_col = (u64)(1u << log_st) >> p.log_a;
        const u64 col = tile / tiles_per_col;
        const u64 k0 = (tile % tiles_per_col) << p.log_a;
        u64 *dst = p.out + col * p.out_col_stride;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 a = idx & (A - 1), k1 = idx >> p.log_a;
            u64 v = gl_mul(sm[ntt_sm(pitch, a, ntt_brev(k1, p.log_p))], p.scale);
            dst[((u64)k1 << log_st) + k0 + a] = gl_canon(v);
        }
    } else {  // NTT_INTT_SINGLE
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 k = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            if (unit < p.num_cols) {
                u64 v = gl_mul(sm[ntt_sm(pitch, a, ntt_brev(k, p.log_p))], p.scale);
                p.out[unit * p.out_col_stride + k] = gl_canon(v);
            }
        }
    }
}

#ifdef __CUDACC__
template <int MODE>
__global__ void __launch_bounds__(512) ntt_pass_kernel(NttPass p) {
    extern __shared__ u64 ntt_smem[];
    for (u64 tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ntt_load<MODE>(p, ntt_smem, tile, threadIdx.x, blockDim.x);
        __syncthreads();
        u32 t0 = 0;
        u32 rem = p.log_p & 3;
        if (rem == 1) { ntt_round<1>(p, ntt_smem, t0, threadIdx.x, blockDim.x); t0 += 1; __syncthreads(); }
        if (rem == 2) { ntt_round<2>(p, ntt_smem, t0, threadIdx.x, blockDim.x); t0 += 2; __syncthreads(); }
        if (rem == 3) { ntt_round<3>(p, ntt_smem, t0, threadIdx.x, blockDim.x); t0 += 3; __syncthreads(); }
        for (; t0 < p.log_p; t0 += 4) { ntt_round<4>(p, ntt_smem, t0, threadIdx.x, blockDim.x); __syncthreads(); }
        ntt_store<MODE>(p, ntt_smem, tile, threadIdx.x, blockDim.x);
        __syncthreads();
    }
}
#endif
