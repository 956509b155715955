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
    // fused exchange: the last pass walks its tiles starting at tile_rot (= this rank's own row shard), so that at any
    // moment the G ranks of the box store into G DIFFERENT destinations instead of all hammering shard 0, then 1, ...
    u64 tile_rot;
    const u64 *tw_local;  // w_P^e (or inverse), e < P
    // tau tables of the non-final radix-16 rounds, round after round: [r - 1][lo] = w_P^(bitrev4(r) * (lo << t0)), r = 1..15,
    // lo < 2^(log_p - t0 - 4).  Consecutive threads hold consecutive lo, so the 15 tau loads of a block are conflict-free;
    // indexing the w_P^e table with bitrev4(r) * (lo << t0) instead strides by up to 256 bytes between threads (ncu, round 2:
    // 429 M of the 495 M shared-memory wavefronts of those loads were bank-conflict replays).
    const u64 *tau_tab;
    u32 tau_in_smem;      // the kernel stages tau_tab in shared memory (it fits); otherwise it is read through L1
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

// Lane pitch.  P + P/16 keeps both access patterns of the rounds conflict-free; the pad decides the strided LOAD phase,
// where the lane index runs fastest over the threads: with A lanes a warp stores (a, j), a < A, j < 32 / A, and
// a * pitch + j must hit 32 different 8-byte slots mod 32: pitch = 32 / A (mod 32) (1 for P < 32).
GL_HD u32 ntt_pitch(u32 log_p, u32 log_a) {
    const u32 P = 1u << log_p, base = P + (P >> 4);
    if (log_p < 5 || log_a == 0 || log_a > 5) return base + 1;
    const u32 want = 32u >> log_a;                       // 32 / A
    return base + ((want + 32u - (base & 31u)) & 31u);
}
GL_HD u32 ntt_sm(u32 pitch, u32 a, u32 j) { return a * pitch + j + (j >> 4); }
// entries of the tau tables of a P-point network (the rounds: a remainder round of log_p mod 4 stages, then radix-16)
GL_HD u32 ntt_tau_entries(u32 log_p) {
    u32 total = 0;
    for (u32 t0 = log_p & 3; t0 + 4 < log_p; t0 += 4) total += 15u << (log_p - t0 - 4);
    return total;
}
#define NTT_SMEM_LIMIT (227u * 1024u)
static inline size_t ntt_smem_bytes_base(u32 log_p, u32 log_a) {  // tile + the w_P^e table (all P powers)
    return ((size_t)ntt_pitch(log_p, log_a) * (1u << log_a) + (1u << log_p) + 1) * sizeof(u64);
}
static inline bool ntt_tau_fits(u32 log_p, u32 log_a) { return ntt_smem_bytes_base(log_p, log_a) + (size_t)ntt_tau_entries(log_p) * 8 <= NTT_SMEM_LIMIT; }
static inline size_t ntt_smem_bytes(u32 log_p, u32 log_a) {
    return ntt_smem_bytes_base(log_p, log_a) + (ntt_tau_fits(log_p, log_a) ? (size_t)ntt_tau_entries(log_p) * 8 : 0);
}

GL_HD u64 ntt_twiddle2(const NttPass &p, u64 e) {
    u64 lo = p.w_lo[e & ((1ull << p.w_lo_bits) - 1)];
    u64 hi = p.w_hi[e >> p.w_lo_bits];
    return gl_mul(lo, hi);
}

// ---- phase 1: global -> shared ----
// Loads are issued NTT_LOAD_BATCH at a time before the first one is consumed (the phase is DRAM/L2 latency).
#ifndef NTT_LOAD_BATCH
#define NTT_LOAD_BATCH 4
#endif
template <int MODE>
GL_HD void ntt_load(const NttPass &p, u64 *sm, u64 tile, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, A = 1u << p.log_a, pitch = ntt_pitch(p.log_p, p.log_a);
    const u32 total = P << p.log_a;
    constexpr int B = NTT_LOAD_BATCH;
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
        const u64 *src = p.in + col * p.in_col_stride + r0;
        const u64 *sh = (MODE == NTT_LDE_FIRST) ? p.shift_a + ((u64)e << p.log_p) : nullptr;
        for (u32 i0 = tid; i0 < total; i0 += nthreads * B) {
            u64 v[B], w[B];
#pragma unroll
            for (int k = 0; k < B; k++) {
                const u32 idx = i0 + (u32)k * nthreads;
                if (idx < total) {
                    const u32 a = idx & (A - 1), j = idx >> p.log_a;
                    v[k] = src[((u64)j << log_st) + a];
                    if (MODE == NTT_LDE_FIRST) w[k] = sh[j];   // s_e^(j*st); s_e^r at the store
                }
            }
#pragma unroll
            for (int k = 0; k < B; k++) {
                const u32 idx = i0 + (u32)k * nthreads;
                if (idx < total) {
                    const u32 a = idx & (A - 1), j = idx >> p.log_a;
                    sm[ntt_sm(pitch, a, j)] = (MODE == NTT_LDE_FIRST) ? gl_mul(v[k], w[k]) : v[k];
                }
            }
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
#ifdef __CUDA_ARCH__
        if (MODE == NTT_DIF_LAST && p.log_p >= 1) {
            // 16-byte accesses: element pairs (j, j+1), j even
            const u32 pairs = total >> 1;
            for (u32 i0 = tid; i0 < pairs; i0 += nthreads * B) {
                ulonglong2 v[B];
#pragma unroll
                for (int k = 0; k < B; k++) {
                    const u32 idx = (i0 + (u32)k * nthreads) << 1;
                    v[k] = make_ulonglong2(0, 0);
                    if (idx < total) {
                        const u64 unit = tile * A + (idx >> p.log_p);
                        if (unit < p.num_units) v[k] = *reinterpret_cast<const ulonglong2 *>(p.in + (unit << p.log_p) + (idx & (P - 1)));
                    }
                }
#pragma unroll
                for (int k = 0; k < B; k++) {
                    const u32 idx = (i0 + (u32)k * nthreads) << 1;
                    if (idx < total) {
                        const u32 a = idx >> p.log_p, j = idx & (P - 1);
                        sm[ntt_sm(pitch, a, j)] = v[k].x;
                        sm[ntt_sm(pitch, a, j + 1)] = v[k].y;
                    }
                }
            }
            return;
        }
#endif
        for (u32 i0 = tid; i0 < total; i0 += nthreads * B) {
            u64 v[B];
#pragma unroll
            for (int k = 0; k < B; k++) {
                const u32 idx = i0 + (u32)k * nthreads;
                v[k] = 0;
                if (idx < total) {
                    const u32 j = idx & (P - 1), a = idx >> p.log_p;
                    const u64 unit = tile * A + a;
                    if (MODE == NTT_DIF_LAST) {
                        // the LDE buffer is one contiguous array of aligned P-point runs, whatever its shard layout
                        if (unit < p.num_units) v[k] = p.in[(unit << p.log_p) + j];
                    } else {
                        if (unit < p.num_cols) v[k] = p.in[unit * p.in_col_stride + j];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < B; k++) {
                const u32 idx = i0 + (u32)k * nthreads;
                if (idx < total) sm[ntt_sm(pitch, idx >> p.log_p, idx & (P - 1))] = v[k];
            }
        }
    }
}

// ---- phase 2: one register-blocked round of R DIF stages (stages t0+1 .. t0+R of the P-point network) ----
// `tw` is the w_P^e table (the kernel stages it in shared memory).  LAST marks the final round (element stride 1):
// there the twiddle of a butterfly depends only on its register index, and the ones that are 1 are skipped.
template <int R, bool LAST>
GL_HD void ntt_round(const NttPass &p, u64 *sm, const u64 *tw, u32 t0, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, pitch = ntt_pitch(p.log_p, p.log_a);
    const u32 blocks_per_lane = P >> R;
    const u32 nblocks = blocks_per_lane << p.log_a;
    const u32 log_js = p.log_p - t0 - R;  // log2 of the element stride inside the register block
    // Addressing is kept off the alu pipe: the 16 slots of a register block are base + m * step (the padding j >> 4 is
    // linear in m whenever the element stride is 1 or a multiple of 16), and a twiddle exponent is
    // (lo << s_u) + m' * 2^(log_js + s_u) -- one IMAD each instead of shift/add/shift/add chains.
    const u32 log_bpl = p.log_p - R;
    const bool linear = (log_js >= 4) || (log_js == 0);
    const u32 step = log_js >= 4 ? ((1u << log_js) + (1u << (log_js - 4))) : 1u;
    for (u32 blk = tid; blk < nblocks; blk += nthreads) {
        u32 a = blk >> log_bpl, T = blk & (blocks_per_lane - 1);
        u32 lo = T & ((1u << log_js) - 1), hi = T >> log_js;
        u32 jb = (hi << (p.log_p - t0)) + lo;
        const u32 base = ntt_sm(pitch, a, jb);
        u64 v[1 << R];
        if (linear) {
#pragma unroll
            for (int m = 0; m < (1 << R); m++) v[m] = sm[base + (u32)m * step];
        } else {
#pragma unroll
            for (int m = 0; m < (1 << R); m++) v[m] = sm[ntt_sm(pitch, a, jb + ((u32)m << log_js))];
        }
#pragma unroll
        for (int u = 1; u <= R; u++) {
            const int span = 1 << (R - u);
            const u32 su = t0 + u - 1;
            const u32 e0 = lo << su, em = 1u << (log_js + su);   // exponent of register m' = e0 + m' * em
#pragma unroll
            for (int m = 0; m < (1 << R); m++) {
                if (m & span) continue;
                // pair (j, j + half), half = span << log_js; twiddle w_P^{(j mod half) << (t0+u-1)}
                u64 x = v[m], y = v[m + span];
                v[m] = gl_add(x, y);
                if (LAST && (m & (span - 1)) == 0) v[m + span] = gl_sub(x, y);  // w = 1
                else v[m + span] = gl_mul(gl_sub(x, y), tw[e0 + (u32)(m & (span - 1)) * em]);
            }
        }
        if (linear) {
#pragma unroll
            for (int m = 0; m < (1 << R); m++) sm[base + (u32)m * step] = v[m];
        } else {
#pragma unroll
            for (int m = 0; m < (1 << R); m++) sm[ntt_sm(pitch, a, jb + ((u32)m << log_js))] = v[m];
        }
    }
}


// ---- radix-16 round as ONE 16-point DFT with power-of-two twiddles + one twiddle product per output ----------------
// In Goldilocks 2 has order 192 and plonky2's roots of unity satisfy w_16 = 2^156, w_8 = 2^120, w_4 = 2^48 (w_16 is
// h_gl_root_of_unity(4); checked by tests/test_replay.py), so every twiddle INSIDE a 16-point DFT is a shift.  The four
// stages of a radix-16 block multiply the lower branch of stage u by w_{2^(5-u)}^{m'} * tau^(2^(u-1)), tau =
// w_P^(lo << t0); the tau factors commute to the outputs: register r leaves the pure DFT times tau^bitrev4(r).  So a
// block is 64 additions / subtractions, 17 shifts and 15 (instead of 32) twiddle products.  The additions run LAZILY on
// 96-bit two's-complement values (3 instructions each, no wrap correction); a value is reduced once, before its product.
struct gl96 {
    u32 w0, w1, w2;   // w0 + 2^32 w1 + 2^64 (int32)w2
};
GL_HD gl96 l3_from(u64 x) { gl96 r; r.w0 = (u32)x; r.w1 = (u32)(x >> 32); r.w2 = 0; return r; }
#ifndef __CUDA_ARCH__
typedef __int128 l3_i128;
GL_HD l3_i128 l3_val(gl96 a) { return (l3_i128)a.w0 + ((l3_i128)a.w1 << 32) + ((l3_i128)(int32_t)a.w2) * ((l3_i128)1 << 64); }
// L3_TRACK: host-only hook of the CPU replay recording the largest lazy magnitude (the 96-bit two's-complement form and
// l3_shl's 128-bit intermediate need |v| < 2^80)
#ifndef L3_TRACK
#define L3_TRACK(v)
#endif
GL_HD gl96 l3_make(l3_i128 v) {
    L3_TRACK(v);
    gl96 r; r.w0 = (u32)v; r.w1 = (u32)(v >> 32); r.w2 = (u32)(v >> 64);
    return r;   // callers stay far below 2^95; the GPU code wraps identically
}
#endif
GL_HD gl96 l3_add(gl96 a, gl96 b) {
#ifdef __CUDA_ARCH__
    gl96 r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=&r"(r.w0), "=&r"(r.w1), "=r"(r.w2) : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
#else
    return l3_make(l3_val(a) + l3_val(b));
#endif
}
GL_HD gl96 l3_sub(gl96 a, gl96 b) {
#ifdef __CUDA_ARCH__
    gl96 r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=&r"(r.w0), "=&r"(r.w1), "=r"(r.w2) : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
#else
    return l3_make(l3_val(a) - l3_val(b));
#endif
}
// x * 2^S (mod p) for a lazy x with |x| < 2^80, 0 < S < 96, S = 32 a + b.  Y = x << b is the 128-bit two's-complement
// number (y0, y1, y2, y3); 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32, 2^160 = 1 - 2^32 (mod p) give
//   a = 0:  (y0 - y2 - y3) + 2^32 (y1 + y2)        a = 1:  (-y1 - y2) + 2^32 (y0 + y1 - y3)
//   a = 2:  (-y0 - y1 + y3) + 2^32 (y0 - y2 - y3)
template <int S>
GL_HD gl96 l3_shl(gl96 x) {
    constexpr int A = S / 32, B = S % 32;
#ifdef __CUDA_ARCH__
    u32 y0, y1, y2, y3;
    if (B == 0) {
        y0 = x.w0; y1 = x.w1; y2 = x.w2; y3 = (u32)((int32_t)x.w2 >> 31);
    } else {
        y0 = x.w0 << B;
        y1 = __funnelshift_l(x.w0, x.w1, B);
        y2 = __funnelshift_l(x.w1, x.w2, B);
        y3 = (u32)((int32_t)x.w2 >> (32 - B));
    }
    const u32 s3 = (u32)((int32_t)y3 >> 31);
    gl96 r;
    if (A == 0) {
        asm("{\n\t"
            "sub.cc.u32 %0, %3, %5;\n\t"      // (y0, y1, 0) - y2
            "subc.cc.u32 %1, %4, 0;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "add.cc.u32 %1, %1, %5;\n\t"      // + y2 << 32
            "addc.u32 %2, %2, 0;\n\t"
            "sub.cc.u32 %0, %0, %6;\n\t"      // - y3 (sign extended)
            "subc.cc.u32 %1, %1, %7;\n\t"
            "subc.u32 %2, %2, %7;\n\t"
            "}"
            : "=&r"(r.w0), "=&r"(r.w1), "=&r"(r.w2) : "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(s3));
    } else if (A == 1) {
        asm("{\n\t"
            "sub.cc.u32 %0, 0, %4;\n\t"       // (0, y0, 0) - y1
            "subc.cc.u32 %1, %3, 0;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %5;\n\t"      // - y2
            "subc.cc.u32 %1, %1, 0;\n\t"
            "subc.u32 %2, %2, 0;\n\t"
            "add.cc.u32 %1, %1, %4;\n\t"      // + y1 << 32
            "addc.u32 %2, %2, 0;\n\t"
            "sub.cc.u32 %1, %1, %6;\n\t"      // - y3 << 32 (sign extended)
            "subc.u32 %2, %2, %7;\n\t"
            "}"
            : "=&r"(r.w0), "=&r"(r.w1), "=&r"(r.w2) : "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(s3));
    } else {
        asm("{\n\t"
            "sub.cc.u32 %0, 0, %3;\n\t"       // (0, y0, 0) - y0
            "subc.cc.u32 %1, %3, 0;\n\t"
            "subc.u32 %2, 0, 0;\n\t"
            "sub.cc.u32 %0, %0, %4;\n\t"      // - y1
            "subc.cc.u32 %1, %1, 0;\n\t"
            "subc.u32 %2, %2, 0;\n\t"
            "sub.cc.u32 %1, %1, %5;\n\t"      // - y2 << 32
            "subc.u32 %2, %2, 0;\n\t"
            "add.cc.u32 %0, %0, %6;\n\t"      // + y3 (sign extended)
            "addc.cc.u32 %1, %1, %7;\n\t"
            "addc.u32 %2, %2, %7;\n\t"
            "sub.cc.u32 %1, %1, %6;\n\t"      // - y3 << 32
            "subc.u32 %2, %2, %7;\n\t"
            "}"
            : "=&r"(r.w0), "=&r"(r.w1), "=&r"(r.w2) : "r"(y0), "r"(y1), "r"(y2), "r"(y3), "r"(s3));
    }
    return r;
#else
    const l3_i128 Y = l3_val(x) * ((l3_i128)1 << B);
    const l3_i128 y0 = (u32)Y, y1 = (u32)(Y >> 32), y2 = (u32)(Y >> 64), y3 = (l3_i128)(int32_t)(u32)(Y >> 96);
    const l3_i128 W = (l3_i128)1 << 32;
    if (A == 0) return l3_make((y0 - y2 - y3) + W * (y1 + y2));
    if (A == 1) return l3_make((-y1 - y2) + W * (y0 + y1 - y3));
    return l3_make((-y0 - y1 + y3) + W * (y0 - y2 - y3));
#endif
}
// some u64 representative of a lazy value with |w2| < 2^30:  (w1:w0) + w2 * (2^32 - 1), one fold of the top limb
GL_HD u64 l3_reduce(gl96 x) {
#ifdef __CUDA_ARCH__
    u32 v0, v1;
    asm("{\n\t"
        ".reg .u32 s2, k, tt, hh;\n\t"
        "shr.s32 s2, %4, 31;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"      // (w0, w1, 0) - w2 (sign extended)
        "subc.cc.u32 %1, %3, s2;\n\t"
        "subc.u32 k, 0, s2;\n\t"
        "add.cc.u32 %1, %1, %4;\n\t"      // + w2 << 32 (sign extended)
        "addc.u32 k, k, s2;\n\t"
        "sub.u32 tt, 0, k;\n\t"           // k in {-1, 0, 1}: add k * (2^32 - 1)
        "shr.s32 hh, k, 1;\n\t"
        "add.cc.u32 %0, %0, tt;\n\t"
        "addc.u32 %1, %1, hh;\n\t"
        "}"
        : "=&r"(v0), "=&r"(v1) : "r"(x.w0), "r"(x.w1), "r"(x.w2));
    return ((u64)v1 << 32) | v0;
#else
    l3_i128 v = l3_val(x) % (l3_i128)GL_P;
    if (v < 0) v += (l3_i128)GL_P;
    return (u64)v;
#endif
}

// exponent of 2 of the stage-u twiddle of register m' inside a 16-point DFT (forward: w_16 = 2^156)
__host__ __device__ constexpr int ntt16_shift(int u, int mp, bool inv) { return (((inv ? 192 - 156 : 156) << (u - 1)) * mp) % 192; }
__host__ __device__ constexpr int ntt_brev4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r & 8) >> 3); }

template <int U, int M, bool INV>
GL_HD void ntt16_pair(gl96 (&x)[16]) {
    constexpr int span = 1 << (4 - U);
    constexpr int S = ntt16_shift(U, M & (span - 1), INV);
    const gl96 a = x[M], b = x[M + span];
    x[M] = l3_add(a, b);
    if (S == 0) x[M + span] = l3_sub(a, b);
    else if (S == 96) x[M + span] = l3_sub(b, a);
    else if (S < 96) x[M + span] = l3_shl<(S % 96 == 0 ? 1 : S % 96)>(l3_sub(a, b));
    else x[M + span] = l3_shl<(S % 96 == 0 ? 1 : S % 96)>(l3_sub(b, a));      // 2^(96 + t) = -2^t
}
template <int U, bool INV>
GL_HD void ntt16_stage(gl96 (&x)[16]) {
    constexpr int span = 1 << (4 - U);
    // the 8 butterflies of stage U, registers m with bit `span` clear
    ntt16_pair<U, (0 / span) * 2 * span + 0 % span, INV>(x);
    ntt16_pair<U, (1 / span) * 2 * span + 1 % span, INV>(x);
    ntt16_pair<U, (2 / span) * 2 * span + 2 % span, INV>(x);
    ntt16_pair<U, (3 / span) * 2 * span + 3 % span, INV>(x);
    ntt16_pair<U, (4 / span) * 2 * span + 4 % span, INV>(x);
    ntt16_pair<U, (5 / span) * 2 * span + 5 % span, INV>(x);
    ntt16_pair<U, (6 / span) * 2 * span + 6 % span, INV>(x);
    ntt16_pair<U, (7 / span) * 2 * span + 7 % span, INV>(x);
}

// One radix-16 round (stages t0+1 .. t0+4).  `tw` holds w_P^e for ALL e < P (inverse powers when INV).
// tau: this round's table (unused when LAST)
template <bool LAST, bool INV>
GL_HD void ntt_round16(const NttPass &p, u64 *sm, const u64 *tau, u32 t0, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, pitch = ntt_pitch(p.log_p, p.log_a);
    const u32 blocks_per_lane = P >> 4, log_bpl = p.log_p - 4;
    const u32 nblocks = blocks_per_lane << p.log_a;
    const u32 log_js = p.log_p - t0 - 4;
    const bool linear = (log_js >= 4) || (log_js == 0);
    const u32 step = log_js >= 4 ? ((1u << log_js) + (1u << (log_js - 4))) : 1u;
    for (u32 blk = tid; blk < nblocks; blk += nthreads) {
        const u32 a = blk >> log_bpl, T = blk & (blocks_per_lane - 1);
        const u32 lo = T & ((1u << log_js) - 1), hi = T >> log_js;
        const u32 jb = (hi << (p.log_p - t0)) + lo;
        const u32 base = ntt_sm(pitch, a, jb);
        gl96 x[16];
#pragma unroll
        for (int m = 0; m < 16; m++) x[m] = l3_from(sm[linear ? base + (u32)m * step : ntt_sm(pitch, a, jb + ((u32)m << log_js))]);
        ntt16_stage<1, INV>(x);
        ntt16_stage<2, INV>(x);
        ntt16_stage<3, INV>(x);
        ntt16_stage<4, INV>(x);
        // register r carries tau^bitrev4(r), tau = w_P^(lo << t0): row r - 1 of the round's table at column lo
#pragma unroll
        for (int r = 0; r < 16; r++) {
            u64 o = l3_reduce(x[r]);
            if (!LAST && r != 0) o = gl_mul(o, tau[((u32)(r - 1) << log_js) + lo]);
            sm[linear ? base + (u32)r * step : ntt_sm(pitch, a, jb + ((u32)r << log_js))] = o;
        }
    }
}

// ---- phase 3: shared -> global; slot q of a lane holds frequency brev(q) ----
template <int MODE>
GL_HD void ntt_store(const NttPass &p, const u64 *sm, u64 tile, u32 tid, u32 nthreads) {
    const u32 P = 1u << p.log_p, A = 1u << p.log_a, pitch = ntt_pitch(p.log_p, p.log_a);
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
#ifdef __CUDA_ARCH__
        if (p.log_p >= 1) {
            for (u32 idx = tid << 1; idx < total; idx += nthreads << 1) {   // 16-byte stores: pairs (q, q+1), q even
                const u32 q = idx & (P - 1), a = idx >> p.log_p;
                const u64 unit = tile * A + a;
                if (unit < p.num_units) {
                    const u64 i0 = unit << p.log_p;
                    u64 *dst = p.num_shard_ptrs ? p.shard_out[i0 / p.shard_stride] + i0 % p.shard_stride : p.out + i0;
                    *reinterpret_cast<ulonglong2 *>(dst + q) =
                        make_ulonglong2(gl_canon(sm[ntt_sm(pitch, a, q)]), gl_canon(sm[ntt_sm(pitch, a, q + 1)]));
                }
            }
            return;
        }
#endif
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
#define NTT_ROUNDS(SYNC, INV)                                                                                        \
    {                                                                                                           \
        u32 t0 = 0;                                                                                             \
        const u32 rem = p.log_p & 3;                                                                            \
        if (rem == 1) { if (p.log_p == 1) ntt_round<1, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<1, false>(p, sm, tw, t0, tid, nthreads); t0 += 1; SYNC; } \
        if (rem == 2) { if (p.log_p == 2) ntt_round<2, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<2, false>(p, sm, tw, t0, tid, nthreads); t0 += 2; SYNC; } \
        if (rem == 3) { if (p.log_p == 3) ntt_round<3, true>(p, sm, tw, t0, tid, nthreads); else ntt_round<3, false>(p, sm, tw, t0, tid, nthreads); t0 += 3; SYNC; } \
        const u64 *tau_r = tau;                                                                                 \
        for (; t0 + 4 < p.log_p; t0 += 4) { ntt_round16<false, INV>(p, sm, tau_r, t0, tid, nthreads); tau_r += 15u << (p.log_p - t0 - 4); SYNC; } \
        if (t0 < p.log_p) { ntt_round16<true, INV>(p, sm, tau_r, t0, tid, nthreads); SYNC; }                    \
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
    u64 *tw_s = ntt_smem + (size_t)ntt_pitch(p.log_p, p.log_a) * (1u << p.log_a);   // w_P^e table staged once per CTA
    u64 *tau_s = tw_s + (1u << p.log_p) + 1;                                          // ... and the tau tables, if they fit
    const u32 tid = threadIdx.x, nthreads = blockDim.x;
    for (u32 i = tid; i < (1u << p.log_p); i += nthreads) tw_s[i] = p.tw_local[i];
    if (p.tau_in_smem)
        for (u32 i = tid, n_tau = ntt_tau_entries(p.log_p); i < n_tau; i += nthreads) tau_s[i] = p.tau_tab[i];
    const u64 *tw = tw_s;
    const u64 *tau = p.tau_in_smem ? tau_s : p.tau_tab;
    for (u64 tile_i = blockIdx.x; tile_i < p.num_tiles; tile_i += gridDim.x) {
        u64 tile = tile_i + p.tile_rot;
        if (tile >= p.num_tiles) tile -= p.num_tiles;
        ntt_load<MODE>(p, sm, tile, tid, nthreads);
        __syncthreads();
        NTT_ROUNDS(__syncthreads(), (MODE >= NTT_INTT_P1))
        ntt_store<MODE>(p, sm, tile, tid, nthreads);
        __syncthreads();
    }
}
#endif
