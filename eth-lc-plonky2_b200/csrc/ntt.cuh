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

#define NTT_MAX_LOGP 13
#define NTT_MAX_SHARDS 16

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
    u64 num_units;     // NTT_DIF_LAST: number of contiguous P-point runs in the buffer
    // LDE output layout: row k of column c lives at out[(k >> log_shard_rows) * shard_stride + c * out_col_stride +
    // (k & (2^log_shard_rows - 1))].  One shard (log_shard_rows = log2 L) is the plain [C][L] layout; with G row
    // shards the buffer is [G][C][L/G], i.e. the all-to-all send chunks of the multi-GPU commit are contiguous.
    u32 log_shard_rows;
    u64 shard_stride;
    // Fused exchange (multi-GPU): when num_shard_ptrs != 0 the LAST pass of the LDE stores row shard g through
    // shard_out[g] instead of out + g * shard_stride -- shard_out[g] is a peer-mapped pointer (NVLink P2P) to where this
    // rank's [C_r][L/G] block lives inside row-shard owner g's [C][L/G] leaf matrix, so the all-to-all is the store.
    u32 num_shard_ptrs;
    u64 *shard_out[NTT_MAX_SHARDS];
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
static inline size_t ntt_smem_bytes(u32 log_p, u32 log_a) {  // tile + the w_P^e table
    return ((size_t)ntt_pitch(log_p) * (1u << log_a) + ((1u << log_p) >> 1) + 1) * sizeof(u64);
}

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
            if (MODE == NTT_LDE_FIRST) v = gl_mul(v, p.shift_a[((u64)e << p.log_p) + j]);  // s_e^(j*st); s_e^r at the store
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
                // the LDE buffer is one contiguous array of aligned P-point runs, whatever its shard layout
                if (unit < p.num_units) v = p.in[(unit << p.log_p) + j];
            } else {
                if (unit < p.num_cols) v = p.in[unit * p.in_col_stride + j];
            }
            sm[ntt_sm(pitch, a, j)] = v;
        }
    }
}

// ---- phase 2: one register-blocked round of R DIF stages (stages t0+1 .. t0+R of the P-point network) ----
// `tw` is the w_P^e table (the kernel stages it in shared memory).  LAST marks the final round (element stride 1):
// there the twiddle of a butterfly depends only on its register index, and the ones that are 1 are skipped.
template <int R, bool LAST>
GL_HD void ntt_round(const NttPass &p, u64 *sm, const u64 *tw, u32 t0, u32 tid, u32 nthreads) {
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
                u64 x = v[m], y = v[m + span];
                v[m] = gl_add(x, y);
                if (LAST && (m & (span - 1)) == 0) v[m + span] = gl_sub(x, y);  // w = 1
                else v[m + span] = gl_mul(gl_sub(x, y), tw[jm << (t0 + u - 1)]);
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
        u64 *dst = p.out + col * p.out_col_stride;
        const u64 shard_mask = ((u64)1 << p.log_shard_rows) - 1;
        // element (q, a) is multiplied by (s_e * w_n^k1)^(r0+a), k1 = brev(q): one thread walks the A lanes of a slot
        // with a running power (1 product per element instead of a two-level lookup + the coset factor)
        const u64 *sb = p.shift_b + ((u64)e << log_st);
        const u64 se_r0 = sb[r0], se = (log_st ? sb[1] : 1);
        for (u32 q = tid; q < P; q += nthreads) {
            u32 k1 = ntt_brev(q, p.log_p);
            u64 t = gl_mul(ntt_twiddle2(p, r0 * (u64)k1), se_r0);
            u64 step = gl_mul(p.w_lo[k1], se);
            u64 k0 = ((u64)b << p.log_n) + ((u64)q << log_st) + r0;   // LDE row of lane 0; the A lanes share its shard
            u64 *d = dst + (k0 >> p.log_shard_rows) * p.shard_stride + (k0 & shard_mask);
            for (u32 a = 0; a < A; a += 2) {
                u64 o0 = gl_mul(sm[ntt_sm(pitch, a, q)], t);
                t = gl_mul(t, step);
                u64 o1 = gl_mul(sm[ntt_sm(pitch, a + 1, q)], t);
                t = gl_mul(t, step);
#ifdef __CUDA_ARCH__
                *reinterpret_cast<ulonglong2 *>(d + a) = make_ulonglong2(o0, o1);
#else
                d[a] = o0; d[a + 1] = o1;
#endif
            }
        }
    } else if (MODE == NTT_LDE_SINGLE) {
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 q = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            u64 col = unit >> p.rate_bits;
            u32 e = (u32)(unit & ((1u << p.rate_bits) - 1));
            u32 b = ntt_brev(e, p.rate_bits);
            u64 k = ((u64)b << p.log_n) + q;
            if (col < p.num_cols) {
                u64 *base = p.num_shard_ptrs ? p.shard_out[k >> p.log_shard_rows] : p.out + (k >> p.log_shard_rows) * p.shard_stride;
                base[col * p.out_col_stride + (k & (((u64)1 << p.log_shard_rows) - 1))] = gl_canon(sm[ntt_sm(pitch, a, q)]);
            }
        }
    } else if (MODE == NTT_DIF_LAST) {
        for (u32 idx = tid; idx < total; idx += nthreads) {
            u32 q = idx & (P - 1), a = idx >> p.log_p;
            u64 unit = tile * A + a;
            if (unit < p.num_units) {
                u64 i0 = unit << p.log_p;   // a run never straddles row shards (shard_stride is a multiple of P)
                u64 *dst = p.num_shard_ptrs ? p.shard_out[i0 / p.shard_stride] + i0 % p.shard_stride : p.out + i0;
                dst[q] = gl_canon(sm[ntt_sm(pitch, a, q)]);
            }
        }
    } else if (MODE == NTT_INTT_P1) {
        // Y[col][(r0+a) * P + k2], k2 = brev(q), times w_n^{-(r0+a) k2}
        const u64 tiles_per_col = (u64)(1u << log_st) >> p.log_a;
        const u64 col = tile / tiles_per_col;
        const u64 r0 = (tile % tiles_per_col) << p.log_a;
        u64 *dst = p.out + col * p.out_col_stride;
        for (u32 k2 = tid; k2 < P; k2 += nthreads) {   // running power of w_n^-k2 along the lanes
            u32 q = ntt_brev(k2, p.log_p);
            u64 t = ntt_twiddle2(p, r0 * (u64)k2);
            u64 step = p.w_lo[k2];
            for (u32 a = 0; a < A; a++) {
                dst[((r0 + a) << p.log_p) + k2] = gl_mul(sm[ntt_sm(pitch, a, q)], t);
                t = gl_mul(t, step);
            }
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

// Runs the rounds of one P-point network: remainder round first, radix-16 rounds after, the last one specialised.
#define NTT_ROUNDS(SYNC)                                                                                        \
    {                                                                                                           \
        u32 t0 = 0;                                                                                             \
        const u32 rem = p.log_p & 3;                                                                            \
        if (rem == 1) { if (p.log_p == 1) ntt_round<1, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<1, false>(p, sm, tw, t0, tid, nthreads); t0 += 1; SYNC; } \
        if (rem == 2) { if (p.log_p == 2) ntt_round<2, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<2, false>(p, sm, tw, t0, tid, nthreads); t0 += 2; SYNC; } \
        if (rem == 3) { if (p.log_p == 3) ntt_round<3, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<3, false>(p, sm, tw, t0, tid, nthreads); t0 += 3; SYNC; } \
        for (; t0 + 4 < p.log_p; t0 += 4) { ntt_round<4, false>(p, sm, tw, t0, tid, nthreads); SYNC; }          \
        if (t0 < p.log_p) { ntt_round<4, true>(p, sm, tw, t0, tid, nthreads); SYNC; }                           \
    }

#ifdef __CUDACC__
template <int MODE>
#ifndef NTT_MINB
#define NTT_MINB 2
#endif
#ifndef NTT_LB_THREADS
#define NTT_LB_THREADS 512
#endif
__global__ void __launch_bounds__(NTT_LB_THREADS, NTT_MINB) ntt_pass_kernel(NttPass p) {
    extern __shared__ u64 ntt_smem[];
    u64 *sm = ntt_smem;
    u64 *tw_s = ntt_smem + (size_t)ntt_pitch(p.log_p) * (1u << p.log_a);   // w_P^e table staged once per CTA
    const u32 tid = threadIdx.x, nthreads = blockDim.x;
    for (u32 i = tid; i < ((1u << p.log_p) >> 1); i += nthreads) tw_s[i] = p.tw_local[i];
    const u64 *tw = tw_s;
    for (u64 tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ntt_load<MODE>(p, sm, tile, tid, nthreads);
        __syncthreads();
        NTT_ROUNDS(__syncthreads())
        ntt_store<MODE>(p, sm, tile, tid, nthreads);
        __syncthreads();
    }
}
#endif
