// Kernels of the prover rows beyond the commit: opening evaluations (a7) and batched FRI (a8).
//
// Replaces plonky2::plonk::proof::OpeningSet::new (evaluate every committed polynomial at zeta / g*zeta),
// plonky2::fri::oracle::PolynomialBatch::prove_openings and plonky2::fri::prover::{fri_committed_trees,
// fri_proof_of_work} (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350), reached from
// /root/reference/eth-lc-plonky2/src/main.rs:230.
//
// B200 design (not plonky2's): plonky2 works in COEFFICIENT space -- sum alpha^j f_j over 2^degree_bits
// coefficients, a sequential Horner scan for the division by (X - z), a size-L extension-field FFT, and one more
// coset FFT per fold round.  Every one of those values is also obtainable point-wise from data that is already in
// HBM: the committed LDE columns.  Here the FRI codeword is built directly in the EVALUATION domain
//     F(x) = sum_i alpha^{k_i} * (sum_j alpha^j f_ij(x) - sum_j alpha^j f_ij(z_i)) / (x - z_i),   x = 7 w_L^{bitrev(p)},
// with one thread per LDE row reading coalesced columns, and each fold round is a 16-point inverse DFT + Horner per
// Merkle leaf (the interpolation the verifier performs).  No scan, no large FFT, the layers are born in the
// bit-reversed order the Merkle leaves need.  Results are identical field elements (exact arithmetic).
#pragma once
#include "gl64.cuh"
#include "poseidon.cuh"

// ---- field inversion: x^(p-2), p-2 = (2^31 - 1) * 2^33 + (2^32 - 1)  (76 products) ----
GL_HD u64 gl_sqr_n(u64 x, int n) {
    for (int i = 0; i < n; i++) x = gl_sqr(x);
    return x;
}
GL_HD u64 gl_inverse(u64 x) {
    u64 e1 = x;
    u64 e2 = gl_mul(gl_sqr(e1), e1);
    u64 e4 = gl_mul(gl_sqr_n(e2, 2), e2);
    u64 e8 = gl_mul(gl_sqr_n(e4, 4), e4);
    u64 e16 = gl_mul(gl_sqr_n(e8, 8), e8);
    u64 e24 = gl_mul(gl_sqr_n(e16, 8), e8);
    u64 e28 = gl_mul(gl_sqr_n(e24, 4), e4);
    u64 e30 = gl_mul(gl_sqr_n(e28, 2), e2);
    u64 e31 = gl_mul(gl_sqr(e30), e1);          // x^(2^31 - 1)
    u64 e32 = gl_mul(gl_sqr(e31), e1);          // x^(2^32 - 1)
    return gl_mul(gl_sqr_n(e31, 33), e32);
}
// 1 / (a + bX) = (a - bX) / (a^2 - 7 b^2)
GL_HD gl2 gl2_inverse(gl2 x) {
    u64 nrm = gl_sub(gl_sqr(x.a), gl_mul(gl_sqr(x.b), 7));
    u64 ni = gl_inverse(nrm);
    return gl2_make(gl_mul(x.a, ni), gl_mul(gl_sub(0, gl_canon(x.b)), ni));
}
GL_HD gl2 gl2_mul_base_add(gl2 acc, gl2 y, u64 c) {  // acc * y + c
    gl2 t = gl2_mul(acc, y);
    t.a = gl_add(t.a, c);
    return t;
}

// ---- a7: evaluate polynomials at an extension point -------------------------------------------------------
// Block b of polynomial c covers coefficients [b*T*S, (b+1)*T*S): thread t runs Horner in y = z^T over the
// interleaved coefficients start + t + i*T (coalesced), the block folds sum_t z^t v_t as a tree with the multipliers
// z^(T/2), z^(T/4), ... and writes one partial per block.
#define EVAL_T 256
struct EvalParams {
    const u64 *coeffs;     // [C][n]
    u64 n;
    u32 seg;               // S: coefficients per thread (power of two), chunk = EVAL_T * S <= n
    gl2 y;                 // z^T
    gl2 zpow[9];           // z^(T/2), z^(T/4), ..., z^1   (log2 T entries)
    gl2 *partials;         // [C][n / chunk]
};
#ifdef __CUDACC__
__global__ void __launch_bounds__(EVAL_T) eval_partial_kernel(EvalParams p) {
    __shared__ u64 sa[EVAL_T], sb[EVAL_T];
    const u32 t = threadIdx.x;
    const u64 chunk = (u64)EVAL_T * p.seg;
    const u64 first = (u64)blockIdx.x * chunk + t;
    const u64 *c = p.coeffs + (u64)blockIdx.y * p.n;
    gl2 acc = gl2_make(0, 0);
    for (int i = (int)p.seg - 1; i >= 0; i--) {
        u64 k = first + (u64)i * EVAL_T;
        acc = gl2_mul_base_add(acc, p.y, k < p.n ? c[k] : 0);   // n < 256: the tail of the only chunk is zero
    }
    sa[t] = acc.a; sb[t] = acc.b;
    __syncthreads();
    int lvl = 0;
    for (u32 half = EVAL_T / 2; half >= 1; half >>= 1, lvl++) {
        if (t < half) {
            gl2 hi = gl2_mul(gl2_make(sa[t + half], sb[t + half]), p.zpow[lvl]);
            sa[t] = gl_add(sa[t], hi.a); sb[t] = gl_add(sb[t], hi.b);
        }
        __syncthreads();
    }
    if (t == 0) p.partials[(u64)blockIdx.y * gridDim.x + blockIdx.x] = gl2_make(sa[0], sb[0]);
}
// out[c] = sum_b partial[c][b] * w^b, w = z^chunk; one block per polynomial, Horner over strided partials then a tree
struct EvalFinalParams {
    const gl2 *partials;   // [C][nb]
    u32 nb;                // power of two
    gl2 wpow_t;            // w^T' where T' = threads actually used = min(nb, EVAL_T)
    gl2 wtree[9];          // w^(T'/2), ..., w^1
    gl2 *out;              // [C]
};
__global__ void __launch_bounds__(EVAL_T) eval_final_kernel(EvalFinalParams p) {
    __shared__ u64 sa[EVAL_T], sb[EVAL_T];
    const u32 t = threadIdx.x, T = blockDim.x;   // T = min(nb, EVAL_T), power of two
    const gl2 *src = p.partials + (u64)blockIdx.x * p.nb;
    gl2 acc = gl2_make(0, 0);
    for (int i = (int)(p.nb / T) - 1; i >= 0; i--) {
        gl2 v = src[(u64)i * T + t];
        acc = gl2_mul(acc, p.wpow_t);
        acc.a = gl_add(acc.a, v.a); acc.b = gl_add(acc.b, v.b);
    }
    sa[t] = acc.a; sb[t] = acc.b;
    __syncthreads();
    int lvl = 0;
    for (u32 half = T / 2; half >= 1; half >>= 1, lvl++) {
        if (t < half) {
            gl2 hi = gl2_mul(gl2_make(sa[t + half], sb[t + half]), p.wtree[lvl]);
            sa[t] = gl_add(sa[t], hi.a); sb[t] = gl_add(sb[t], hi.b);
        }
        __syncthreads();
    }
    if (t == 0) p.out[blockIdx.x] = gl2_canon(gl2_make(sa[0], sb[0]));
}
#endif

// ---- a8: FRI ----------------------------------------------------------------------------------------------
#define FRI_MAX_BATCHES 4
struct FriCombineParams {
    u32 log_l;
    u32 num_batches;
    u32 first[FRI_MAX_BATCHES + 1];   // polynomials of batch i are cols[first[i] .. first[i+1])
    const u64 *const *cols;           // device array: pointer to the LDE column (bit-reversed rows) of each polynomial
    const gl2 *alpha_pow;             // alpha^j, j < max batch size
    gl2 reduced[FRI_MAX_BATCHES];     // sum_j alpha^j f_ij(z_i)
    gl2 point[FRI_MAX_BATCHES];       // z_i
    gl2 shift[FRI_MAX_BATCHES];       // alpha^(size of batch i): final = final * shift[i] + quotient_i
    const u64 *w_lo, *w_hi;           // two-level powers of w_L
    u32 w_lo_bits;
    u64 *out;                         // layer 0: [count][2], entry t <-> position p = pos0 + t <-> x = 7 * w_L^{bitrev(p)}
    u64 pos0, count;                  // row shard of the multi-GPU prover (single GPU: 0, L); cols[k] is indexed by p - pos0
};
#ifdef __CUDACC__
__global__ void __launch_bounds__(256) fri_combine_kernel(FriCombineParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    const u64 pos = p.pos0 + t;
    const u64 j = __brevll(pos) >> (64 - p.log_l);
    const u64 x = gl_mul(gl_mul(p.w_lo[j & ((1ull << p.w_lo_bits) - 1)], p.w_hi[j >> p.w_lo_bits]), 7);
    gl2 fin = gl2_make(0, 0);
    for (u32 b = 0; b < p.num_batches; b++) {
        u64 aa = 0, ab = 0;  // sum alpha^j f_j(x): accumulate the two coordinates separately
        for (u32 k = p.first[b]; k < p.first[b + 1]; k++) {
            u64 v = p.cols[k][t];
            gl2 a = p.alpha_pow[k - p.first[b]];
            aa = gl_mul_add(a.a, v, aa);
            ab = gl_mul_add(a.b, v, ab);
        }
        gl2 num = gl2_sub(gl2_make(aa, ab), p.reduced[b]);
        gl2 den = gl2_make(gl_sub(x, p.point[b].a), gl_sub(0, p.point[b].b));
        gl2 q = gl2_mul(num, gl2_inverse(den));
        fin = gl2_add(gl2_mul(fin, p.shift[b]), q);
    }
    fin = gl2_canon(fin);
    reinterpret_cast<ulonglong2 *>(p.out)[t] = make_ulonglong2(fin.a, fin.b);
}

// One thread per Merkle leaf (16 extension values e_i at x0 * w_16^{bitrev4(i)}): a_t = (1/16) sum_m w_16^{-mt} E_m with
// E_m = e_{bitrev4(m)}, then P'(y) = sum_t (beta / x0)^t a_t  (fold of P(x) = sum_t x^t P_t(x^16) at beta).
struct FriFoldParams {
    const u64 *in;        // [N][2], bit-reversed order
    u64 *out;             // [N/16][2]
    u32 log_n;            // log2 N
    u64 w16_inv[16];      // w_16^{-k}
    u64 inv16;            // 1/16
    u64 shift_inv;        // 1 / s_k   (coset shift of this layer)
    gl2 beta;
    const u64 *wi_lo, *wi_hi;  // two-level powers of w_N^{-1}
    u32 wi_lo_bits;
};
__global__ void __launch_bounds__(128) fri_fold_kernel(FriFoldParams p) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 log_c = p.log_n - 4;
    if (c >> log_c) return;
    gl2 E[16];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(p.in) + 16 * c;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        ulonglong2 v = src[i];
        const int m = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
        E[m] = gl2_make(v.x, v.y);
    }
    // x0^{-1} = s^{-1} * w_N^{-j'},  j' = bitrev_{log_c}(c)
    const u64 jp = log_c ? (__brevll(c) >> (64 - log_c)) : 0;
    const u64 x0_inv = gl_mul(gl_mul(p.wi_lo[jp & ((1ull << p.wi_lo_bits) - 1)], p.wi_hi[jp >> p.wi_lo_bits]), p.shift_inv);
    const gl2 gam = gl2_scale(p.beta, x0_inv);
    gl2 acc = gl2_make(0, 0);
#pragma unroll 1
    for (int t = 15; t >= 0; t--) {
        u64 aa = 0, ab = 0;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            u64 w = p.w16_inv[(m * t) & 15];
            aa = gl_mul_add(E[m].a, w, aa);
            ab = gl_mul_add(E[m].b, w, ab);
        }
        acc = gl2_mul(acc, gam);
        acc.a = gl_add(acc.a, aa); acc.b = gl_add(acc.b, ab);
    }
    acc = gl2_canon(gl2_scale(acc, p.inv16));
    reinterpret_cast<ulonglong2 *>(p.out)[c] = make_ulonglong2(acc.a, acc.b);
}

// Proof of work: smallest w in [base, base + count) whose duplex response has `pow_bits` leading zeros.
__global__ void __launch_bounds__(128) fri_pow_kernel(const u64 *state12, u32 pos, u64 base, u64 count, u32 pow_bits, unsigned long long *best) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = state12[k];
#pragma unroll
    for (int k = 0; k < 12; k++) if ((u32)k == pos) s[k] = base + i;
    poseidon_permute(s);
    if ((gl_canon(s[7]) >> (64 - pow_bits)) == 0) atomicMin(best, (unsigned long long)(base + i));
}
#endif
