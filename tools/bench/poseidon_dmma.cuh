// Leaf hashing with the linear layers of Poseidon on the FP64 TENSOR pipe (mma.sync.m8n8k4.f64, "DMMA").
//
// Same function as merkle_leaves_kernel (merkle.cuh) / poseidon_permute_f64 (poseidon_f64.cuh): plonky2's Poseidon sponge
// over Goldilocks (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350).  All arithmetic of the
// linear layers is exact integer arithmetic in binary64 (every partial sum < 2^51), so the digests are bit-identical.
//
// Why (profiles/r02_pipe_model.md): a DMMA occupies the FP64 pipe for 16 cycles per warp = 64 FMA/clk/SM, the DFMA rate,
// but it takes ONE issue slot for 256 multiply-adds, and the integer pipes issue underneath it.  The scalar kernel is bound
// by the issue port (DFMA holds it ~2.2 cycles); here the 19 dense layers of a permutation leave the port.
//
// Layout.  A warp owns 32 sponge states = 4 batches x 8 states.  State (b, g) is spread over the four threads
// lane = 4 g + t, t = 0..3, three lanes each ("slots" c = 0..2): slot c of thread t in batch b holds state lane
//     l(b, c, t) = (4 c + t + b) mod 12,        i.e. "virtual" lane v = 4 c + t, rotated by the batch number.
// The fragments of mma.m8n8k4 (A: thread (g,t) holds A[g][t];  B: B[t][g];  C/D: D[g][2t], D[g][2t+1]) then make
//   * slot c of a batch, across the warp, the A operand of k-chunk c (rows = the 8 states);
//   * D of n-tile 0 the new slots 0 and 1, D of n-tile 1 the new slot 2 and one spare column per thread;
//   * the B operands constants of the thread: the MDS matrix is circulant, so the rotation by b cancels and ONE set of
//     six B registers serves all batches (the + 8 on M[0][0] is a local FMA of the thread that holds lane 0).
// The rotation puts lane 0 of batch b into thread (4 - b) mod 4, so in the partial rounds (one S-box per state and round)
// every thread runs exactly one S-box per round: the one of "its" batch.  The pre-S-box value of the second round of a
// pair, T0 = (M W + RC)[0], arrives through the spare column of n-tile 1 (row (-b) mod 12 of circ in column 2t+1), and
// the rank-one term of the pair, (8 W_0 + d) * C e_0, is broadcast inside the group of four with shuffles.
// Algebra of the rounds: poseidon_f64.cuh (two partial rounds per step, M^2 = C^2 + 8(C E + E C) + 64 E).
#pragma once
#include "../../eth-lc-plonky2_b200/csrc/merkle.cuh"

struct PsdDmmaTables {
    double full_init[8][2][12];   // chain heads of the 8 full-round layers [layer][half][lane]
    double pair_k[11][2][12];     // M RC_{t+1} + RC_{t+2} - B M e_0 (- B where folded next); lane 0 also - 8 pair_t0
    double pair_t0[11][2];        // RC_{t+1}[0] - B
    double circ[12];              // C[r][j] = circ[(j - r) mod 12]
    double circ2[12];             // C^2 likewise
    double c_col0[12];            // (C e_0)[r] = circ[(12 - r) mod 12]
    u64 rc0[12];                  // RC_0
};

#ifdef __CUDACC__
__device__ PsdDmmaTables g_pd;

static inline cudaError_t psd_dmma_upload_tables() {
    static PsdF64Tables f;
    static PsdDmmaTables t;
    psd_f64_build_tables(f);
    for (int L = 0; L < 8; L++)
        for (int h = 0; h < 2; h++)
            for (int j = 0; j < 12; j++) t.full_init[L][h][j] = f.full_init[L][j][h];
    for (int p = 0; p < 11; p++)
        for (int h = 0; h < 2; h++) {
            for (int j = 0; j < 12; j++) t.pair_k[p][h][j] = f.pair_k[p][j][h];
            t.pair_k[p][h][0] -= 8.0 * f.pair_t0[p][h];
            t.pair_t0[p][h] = f.pair_t0[p][h];
        }
    const u64 circ[12] = POSEIDON_MDS_CIRC_INIT;
    for (int m = 0; m < 12; m++) {
        u64 a = 0;
        for (int i = 0; i < 12; i++) a += circ[i] * circ[(m - i + 12) % 12];
        t.circ[m] = (double)circ[m];
        t.circ2[m] = (double)a;
        t.c_col0[m] = (double)circ[(12 - m) % 12];
        t.rc0[m] = POSEIDON_RC[m];
    }
    return cudaMemcpyToSymbol(g_pd, &t, sizeof(t));
}

__device__ __forceinline__ void pd_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ u32 pd_mod12(u32 v) { return v >= 12 ? v - 12 : v; }   // v < 24
__device__ __forceinline__ u32 pd_lane(u32 b, u32 c, u32 t) { return pd_mod12(4 * c + t + b); }
__device__ __forceinline__ bool pd_is0(u32 b, u32 c, u32 t) { return b == 0 ? (c == 0 && t == 0) : (c == 2 && t + b == 4); }

// B operands of thread (g, t) for y = circ(cc) x:  B[k = t][n = g] of k-chunk c and n-tile 0 / 1.
// n-tile 0, column 2t'+e -> virtual lane 4e + t';  n-tile 1, column 2t' -> virtual lane 8 + t', column 2t'+1 -> spare:
// with `extra`, row (0, 9, 10, 11)[t'] of circ(circ) = the row of real lane 0 in the batch whose lane 0 thread t' holds.
__device__ __forceinline__ void pd_load_b(const double *cc, const double *circ, bool extra, u32 t, u32 g, double (&b0)[3], double (&b1)[3]) {
    const u32 tp = g >> 1, e = g & 1;
#pragma unroll
    for (u32 c = 0; c < 3; c++) {
        const u32 vin = 4 * c + t;
        b0[c] = cc[pd_mod12(vin + 12 - (4 * e + tp))];
        if (e == 0) b1[c] = cc[pd_mod12(vin + 4 - tp)];
        else b1[c] = extra ? circ[pd_mod12(vin + 12 - (tp ? 8 + tp : 0))] : 0.0;
    }
}

// One full round on integer slots that already hold state + RC_t:  S-box, MDS (+ RC_{t+1} - B), fold.
__device__ __forceinline__ void pd_full_round(u64 (&s)[4][3], int L, const PsdDmmaTables &T, const double (&b0)[3], const double (&b1)[3], u32 t) {
#pragma unroll
    for (u32 b = 0; b < 4; b++) {
        double xl[3], xh[3];
#pragma unroll
        for (u32 c = 0; c < 3; c++) pf_pow7(s[b][c], xl[c], xh[c]);
        const u32 l0 = pd_lane(b, 0, t), l1 = pd_lane(b, 1, t), l2 = pd_lane(b, 2, t);
        double d00 = T.full_init[L][0][l0], d01 = T.full_init[L][0][l1], d10 = T.full_init[L][0][l2], d11 = 0.0;
        double e00 = T.full_init[L][1][l0], e01 = T.full_init[L][1][l1], e10 = T.full_init[L][1][l2], e11 = 0.0;
#pragma unroll
        for (u32 c = 0; c < 3; c++) {
            pd_dmma(d00, d01, xl[c], b0[c]);
            pd_dmma(d10, d11, xl[c], b1[c]);
            pd_dmma(e00, e01, xh[c], b0[c]);
            pd_dmma(e10, e11, xh[c], b1[c]);
        }
        if (pd_is0(b, 0, t)) { d00 = fma(xl[0], 8.0, d00); e00 = fma(xh[0], 8.0, e00); }   // M[0][0] = circ[0] + 8
        if (pd_is0(b, 2, t)) { d10 = fma(xl[2], 8.0, d10); e10 = fma(xh[2], 8.0, e10); }
        s[b][0] = pf_fold(d00, e00);
        s[b][1] = pf_fold(d01, e01);
        s[b][2] = pf_fold(d10, e10);
    }
}

// value of the slot that holds lane 0 of "my" batch (thread t: batch (4 - t) mod 4; slot 0 for batch 0, else slot 2)
#define PD_PICK0(dst, arr, t) do { dst = arr[0][0]; if (t == 3) dst = arr[1][2]; if (t == 2) dst = arr[2][2]; if (t == 1) dst = arr[3][2]; } while (0)

// The 22 partial rounds, two per step.  In: integer slots + RC_4.  Out: integer slots + RC_26.
__device__ __forceinline__ void pd_partial_rounds(u64 (&s)[4][3], const PsdDmmaTables &T, u32 lane) {
    const u32 t = lane & 3, g = lane >> 2;
    double al[4][3], ah[4][3];
#pragma unroll
    for (u32 b = 0; b < 4; b++)
#pragma unroll
        for (u32 c = 0; c < 3; c++) {
            al[b][c] = pf_cvt((u32)s[b][c]);
            ah[b][c] = pf_cvt((u32)(s[b][c] >> 32));
            if (pd_is0(b, c, t)) { al[b][c] -= 2251799813685248.0; ah[b][c] -= 2251799813685248.0; }   // enters through a fold
        }
    double b0[3], b1[3];
    pd_load_b(T.circ2, T.circ, true, t, g, b0, b1);
#pragma unroll 1
    for (int p = 0; p < 11; p++) {
        double xl, xh;
        PD_PICK0(xl, al, t);
        PD_PICK0(xh, ah, t);
        const u64 a = pf_fold(xl, xh);
#pragma unroll
        for (u32 b = 0; b < 4; b++)
#pragma unroll
            for (u32 c = 0; c < 3; c++) pf_renorm(al[b][c], ah[b][c]);
        double w0l, w0h;
        pf_pow7(a, w0l, w0h);                            // W_0
        if (t == 0) { al[0][0] = w0l; ah[0][0] = w0h; }
        if (t == 3) { al[1][2] = w0l; ah[1][2] = w0h; }
        if (t == 2) { al[2][2] = w0l; ah[2][2] = w0h; }
        if (t == 1) { al[3][2] = w0l; ah[3][2] = w0h; }
        double el[4][1], eh[4][1];
#pragma unroll
        for (u32 b = 0; b < 4; b++) {                    // C^2 W + K  and, in the spare column, (C W)[0] + pair_t0
            const u32 l0 = pd_lane(b, 0, t), l1 = pd_lane(b, 1, t), l2 = pd_lane(b, 2, t);
            double d00 = T.pair_k[p][0][l0], d01 = T.pair_k[p][0][l1], d10 = T.pair_k[p][0][l2], d11 = T.pair_t0[p][0];
            double e00 = T.pair_k[p][1][l0], e01 = T.pair_k[p][1][l1], e10 = T.pair_k[p][1][l2], e11 = T.pair_t0[p][1];
#pragma unroll
            for (u32 c = 0; c < 3; c++) {
                pd_dmma(d00, d01, al[b][c], b0[c]);
                pd_dmma(d10, d11, al[b][c], b1[c]);
                pd_dmma(e00, e01, ah[b][c], b0[c]);
                pd_dmma(e10, e11, ah[b][c], b1[c]);
            }
            al[b][0] = d00; al[b][1] = d01; al[b][2] = d10; el[b][0] = d11;
            ah[b][0] = e00; ah[b][1] = e01; ah[b][2] = e10; eh[b][0] = e11;
        }
        double t0l = el[0][0], t0h = eh[0][0];
        if (t == 3) { t0l = el[1][0]; t0h = eh[1][0]; }
        if (t == 2) { t0l = el[2][0]; t0h = eh[2][0]; }
        if (t == 1) { t0l = el[3][0]; t0h = eh[3][0]; }
        t0l = fma(w0l, 8.0, t0l);                        // T0 = (C W)[0] + 8 W_0 + pair_t0
        t0h = fma(w0h, 8.0, t0h);
        const u64 bb = pf_fold(t0l, t0h);                // lane 0 entering the second S-box
        double bl, bh;
        pf_pow7(bb, bl, bh);
        const double ul = fma(w0l, 8.0, bl - t0l), uh = fma(w0h, 8.0, bh - t0h);   // 8 W_0 + d,  d = (b' - b) + B
#pragma unroll
        for (u32 b = 0; b < 4; b++) {
            const int src = (int)((lane & ~3u) | ((4 - b) & 3));
            const double vl = __shfl_sync(0xffffffffu, ul, src), vh = __shfl_sync(0xffffffffu, uh, src);
#pragma unroll
            for (u32 c = 0; c < 3; c++) {
                const double m = T.c_col0[pd_lane(b, c, t)];
                al[b][c] = fma(m, vl, al[b][c]);
                ah[b][c] = fma(m, vh, ah[b][c]);
            }
        }
        // lane 0:  8 (C W)[0] + 64 W_0 + 8 d = 8 (T0 + d) = 8 b'   (the constants are inside pair_k)
        if (t == 0) { al[0][0] = fma(bl, 8.0, al[0][0]); ah[0][0] = fma(bh, 8.0, ah[0][0]); }
        if (t == 3) { al[1][2] = fma(bl, 8.0, al[1][2]); ah[1][2] = fma(bh, 8.0, ah[1][2]); }
        if (t == 2) { al[2][2] = fma(bl, 8.0, al[2][2]); ah[2][2] = fma(bh, 8.0, ah[2][2]); }
        if (t == 1) { al[3][2] = fma(bl, 8.0, al[3][2]); ah[3][2] = fma(bh, 8.0, ah[3][2]); }
    }
#pragma unroll
    for (u32 b = 0; b < 4; b++)
#pragma unroll
        for (u32 c = 0; c < 3; c++) s[b][c] = pf_fold(al[b][c], ah[b][c]);
}

// The permutation of the warp's 32 states.
__device__ __forceinline__ void pd_permute(u64 (&s)[4][3], const PsdDmmaTables &T, u32 lane) {
    const u32 t = lane & 3, g = lane >> 2;
#pragma unroll
    for (u32 b = 0; b < 4; b++)
#pragma unroll
        for (u32 c = 0; c < 3; c++) s[b][c] = gl_add_c(s[b][c], T.rc0[pd_lane(b, c, t)]);
    double b0[3], b1[3];
    pd_load_b(T.circ, T.circ, false, t, g, b0, b1);
#pragma unroll 1
    for (int L = 0; L < 8; L++) {
        if (L == 4) {
            pd_partial_rounds(s, T, lane);
            pd_load_b(T.circ, T.circ, false, t, g, b0, b1);
        }
        pd_full_round(s, L, T, b0, b1, t);
    }
}

#ifndef PD_MINB
#define PD_MINB 4
#endif
// merkle_leaves_kernel for leaves of more than noop_max elements and num_leaves a multiple of 32.
// State (b, g) of warp w is leaf 32 w + 8 b + g: a load of one slot touches 4 columns x 8 consecutive rows.
__global__ void __launch_bounds__(128, PD_MINB) merkle_leaves_dmma_kernel(MerkleParams p) {
    __shared__ PsdDmmaTables T;
    {
        const u64 *src = reinterpret_cast<const u64 *>(&g_pd);
        u64 *dst = reinterpret_cast<u64 *>(&T);
        for (u32 i = threadIdx.x; i < sizeof(PsdDmmaTables) / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    const u64 row0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;
    if (row0 >= p.num_leaves) return;   // warp-uniform
    u64 s[4][3];
#pragma unroll
    for (u32 b = 0; b < 4; b++)
#pragma unroll
        for (u32 c = 0; c < 3; c++) s[b][c] = 0;
    const u32 chunks = (p.width + 7) / 8;
#pragma unroll 1
    for (u32 ch = 0; ch < chunks; ch++) {
#pragma unroll
        for (u32 b = 0; b < 4; b++) {
            const u64 *row = p.data + (row0 + 8 * b + g) * p.row_stride;
#pragma unroll
            for (u32 c = 0; c < 3; c++) {
                const u32 ln = pd_lane(b, c, t), col = 8 * ch + ln;
                if (ln < 8 && col < p.width) s[b][c] = row[(u64)col * p.col_stride];
            }
        }
        pd_permute(s, T, lane);
    }
#pragma unroll
    for (u32 b = 0; b < 4; b++) {
        const u64 j = row0 + 8 * b + g;
        u64 *dst = (p.num_layers == 0) ? p.cap + 4 * j : p.digests + 4 * merkle_digest_pos(p.num_layers, 0, j);
#pragma unroll
        for (u32 c = 0; c < 3; c++) {
            const u32 ln = pd_lane(b, c, t);
            if (ln < 4) dst[ln] = gl_canon(s[b][c]);
        }
    }
}
#endif
