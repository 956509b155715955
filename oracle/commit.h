/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  PARITY UNPINNED: restated from the published
 * algorithm of plonky2 0.1.4 (source absent); cross-checked against the derived vectors of SURVEY.md Appendix C.
 *
 * FFT conventions, PolynomialBatch and MerkleTree:
 *   [DEP plonky2_field:fft.rs, polynomial/mod.rs]           
 *   [DEP plonky2:fri/oracle.rs::PolynomialBatch]             
 *   [DEP plonky2:hash/merkle_tree.rs, hash/merkle_proofs.rs]
 *   [DEP plonky2_util: transpose, reverse_index_bits_in_place]
 * Reached from the reference at /root/reference/eth-lc-plonky2/src/main.rs:227 (build) and :230 (prove).
 */
#ifndef ORACLE_COMMIT_H
#define ORACLE_COMMIT_H
#include <vector>
#include <cstring>
#include <cassert>
#include "gl.h"
#include "poseidon.h"

typedef std::vector<u64> vec64;

/* FftRootTable: per layer lg_m, the powers of primitive_root_of_unity(lg_m) (first half) -- a pure cache. */
static inline const u64 *orc_root_layer(int s) {
    static std::vector<u64> table[GL_TWO_ADICITY + 1];
    std::vector<u64> &t = table[s];
    if (t.empty()) {
        #pragma omp critical(orc_root_table)
        if (t.empty()) {
            size_t h = (size_t)1 << (s - 1);
            std::vector<u64> tmp(h);
            u64 wm = gl_root_of_unity(s);
            tmp[0] = 1;
            for (size_t j = 1; j < h; j++) tmp[j] = gl_mul(tmp[j - 1], wm);
            t.swap(tmp);
        }
    }
    return t.data();
}

/* fft(): values[i] = sum_k c_k w^{ik}, natural order in and out (plonky2: bit-reverse, then DIT layers). */
static inline void orc_fft_inplace(u64 *a, int log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, log_n);
        if (i < j) { u64 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int s = 1; s <= log_n; s++) {
        size_t m = (size_t)1 << s, h = m >> 1;
        const u64 *tw = orc_root_layer(s);
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < h; j++) {
                u64 u = a[k + j], t = gl_mul(a[k + j + h], tw[j]);
                a[k + j] = gl_add(u, t);
                a[k + j + h] = gl_sub(u, t);
            }
    }
}

/* ifft(): forward FFT, then out[0]=v[0]/n, out[n/2]=v[n/2]/n, out[i] <-> out[n-i] scaled by 1/n. */
static inline void orc_ifft_inplace(u64 *a, int log_n) {
    size_t n = (size_t)1 << log_n;
    orc_fft_inplace(a, log_n);
    u64 ninv = gl_inv((u64)n % GL_P);
    a[0] = gl_mul(a[0], ninv);
    if (n > 1) a[n / 2] = gl_mul(a[n / 2], ninv);
    for (size_t i = 1; i < n / 2; i++) {
        u64 x = gl_mul(a[i], ninv), y = gl_mul(a[n - i], ninv);
        a[i] = y; a[n - i] = x;
    }
}

/* coeffs.lde(rate_bits).coset_fft(shift): zero-pad to n*2^r, multiply c_k by shift^k, fft.  Natural order. */
static inline void orc_coset_lde(const u64 *coeffs, int log_n, int rate_bits, u64 shift, u64 *out) {
    size_t n = (size_t)1 << log_n, L = n << rate_bits;
    u64 pw = 1;
    for (size_t k = 0; k < n; k++) { out[k] = gl_mul(gl_canon(coeffs[k]), pw); pw = gl_mul(pw, shift); }
    for (size_t k = n; k < L; k++) out[k] = 0;
    orc_fft_inplace(out, log_n + rate_bits);
}

/* coset_ifft(shift): ifft, then multiply c_k by shift^{-k}. */
static inline void orc_coset_ifft_inplace(u64 *a, int log_n, u64 shift) {
    orc_ifft_inplace(a, log_n);
    u64 si = gl_inv(shift), pw = 1;
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) { a[k] = gl_mul(a[k], pw); pw = gl_mul(pw, si); }
}

/* ---------------- MerkleTree ---------------- */
struct OrcMerkleTree {
    size_t num_leaves = 0, leaf_len = 0;
    int cap_height = 0;
    vec64 leaves;   /* [num_leaves][leaf_len] row-major */
    vec64 digests;  /* [2*(L - 2^h)][4], plonky2's interleaved layout */
    vec64 cap;      /* [2^h][4] */
};

/* fill_subtree: [left region][left child digest][right child digest][right region]; returns the root digest. */
static void orc_fill_subtree(u64 *digests_buf, size_t buf_len, const u64 *leaves, size_t n_leaves, size_t leaf_len,
                             u64 out[4]) {
    assert(n_leaves == buf_len / 2 + 1);
    if (buf_len == 0) { orc_hash_or_noop(leaves, leaf_len, out); return; }
    size_t half = buf_len / 2;
    u64 *left_buf = digests_buf, *left_mem = digests_buf + 4 * (half - 1);
    u64 *right_mem = digests_buf + 4 * half, *right_buf = digests_buf + 4 * (half + 1);
    u64 l[4], r[4];
    if (n_leaves >= 1024) {   /* rayon::join */
        #pragma omp task shared(l) firstprivate(left_buf, half, leaves, n_leaves, leaf_len)
        orc_fill_subtree(left_buf, half - 1, leaves, n_leaves / 2, leaf_len, l);
        #pragma omp task shared(r) firstprivate(right_buf, half, leaves, n_leaves, leaf_len)
        orc_fill_subtree(right_buf, half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
        #pragma omp taskwait
    } else {
        orc_fill_subtree(left_buf, half - 1, leaves, n_leaves / 2, leaf_len, l);
        orc_fill_subtree(right_buf, half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
    }
    memcpy(left_mem, l, 32); memcpy(right_mem, r, 32);
    orc_two_to_one(l, r, out);
}

/* MerkleTree::new(leaves, cap_height); returns 0, or -1 when cap_height > log2(leaves) (plonky2 panics). */
static int orc_merkle_build(OrcMerkleTree &t) {
    size_t L = t.num_leaves;
    int log_l = 0; while (((size_t)1 << log_l) < L) log_l++;
    if (((size_t)1 << log_l) != L || t.cap_height > log_l) return -1;
    size_t ncap = (size_t)1 << t.cap_height;
    size_t num_digests = 2 * (L - ncap);
    t.digests.assign(num_digests * 4, 0);
    t.cap.assign(ncap * 4, 0);
    size_t sub_leaves = L / ncap, sub_digests = num_digests / ncap;
    #pragma omp parallel
    #pragma omp single
    for (size_t s = 0; s < ncap; s++) {
        #pragma omp task firstprivate(s)
        orc_fill_subtree(t.digests.data() + 4 * s * sub_digests, sub_digests,
                         t.leaves.data() + s * sub_leaves * t.leaf_len, sub_leaves, t.leaf_len, &t.cap[4 * s]);
    }
    return 0;
}

/* MerkleTree::prove(leaf_index) -> siblings, bottom-up; returns the number of siblings. */
static inline int orc_merkle_prove(const OrcMerkleTree &t, size_t leaf_index, u64 *siblings) {
    int log_l = 0; while (((size_t)1 << log_l) < t.num_leaves) log_l++;
    int num_layers = log_l - t.cap_height;
    size_t sub_size = ((size_t)1 << (num_layers + 1)) - 2;
    size_t sub_idx = leaf_index >> num_layers;
    const u64 *buf = t.digests.data() + 4 * sub_size * sub_idx;
    size_t pair = leaf_index & (((size_t)1 << num_layers) - 1);
    for (int i = 0; i < num_layers; i++) {
        size_t parity = pair & 1; pair >>= 1;
        size_t sib = 2 * ((pair << (i + 1)) + ((size_t)1 << i) - 1) + (1 - parity);
        memcpy(siblings + 4 * i, buf + 4 * sib, 32);
    }
    return num_layers;
}

/* verify_merkle_proof_to_cap [merkle_proofs.rs]: fold upward, bit 0 => current is the left child. */
static inline int orc_merkle_verify(const u64 *leaf, size_t leaf_len, size_t leaf_index, const u64 *cap,
                                    const u64 *siblings, int num_siblings) {
    u64 cur[4], nxt[4];
    orc_hash_or_noop(leaf, leaf_len, cur);
    size_t idx = leaf_index;
    for (int i = 0; i < num_siblings; i++) {
        if (idx & 1) orc_two_to_one(siblings + 4 * i, cur, nxt); else orc_two_to_one(cur, siblings + 4 * i, nxt);
        memcpy(cur, nxt, 32); idx >>= 1;
    }
    return memcmp(cur, cap + 4 * idx, 32) == 0;
}

/* ---------------- PolynomialBatch ---------------- */
struct OrcBatch {
    int num_polys = 0, degree_log = 0, rate_bits = 0;
    vec64 coeffs;        /* [C][n] */
    OrcMerkleTree tree;  /* leaves [L][C]: leaves[k] = evaluations at 7*w_L^{bitrev(k)} */
    double t_ifft = 0, t_lde = 0, t_transpose = 0, t_tree = 0;  /* stage seconds (plonky2's timed! labels) */
};

double orc_now();

/* from_coeffs: lde_values = coeffs.lde(r).coset_fft(7); transpose; reverse_index_bits; MerkleTree::new. */
static int orc_batch_from_coeffs(OrcBatch &b, int cap_height) {
    int C = b.num_polys, log_n = b.degree_log, r = b.rate_bits, log_l = log_n + r;
    size_t n = (size_t)1 << log_n, L = n << r;
    double t0 = orc_now();
    vec64 lde((size_t)C * L);
    #pragma omp parallel for schedule(dynamic)
    for (int c = 0; c < C; c++) orc_coset_lde(&b.coeffs[(size_t)c * n], log_n, r, GL_GENERATOR, &lde[(size_t)c * L]);
    double t1 = orc_now();
    b.tree.num_leaves = L; b.tree.leaf_len = C; b.tree.cap_height = cap_height;
    b.tree.leaves.resize(L * (size_t)C);
    #pragma omp parallel for schedule(static)
    for (size_t k = 0; k < L; k++) {
        size_t j = bitrev(k, log_l);
        u64 *row = &b.tree.leaves[k * C];
        for (int c = 0; c < C; c++) row[c] = lde[(size_t)c * L + j];
    }
    vec64().swap(lde);
    double t2 = orc_now();
    int rc = orc_merkle_build(b.tree);
    double t3 = orc_now();
    b.t_lde = t1 - t0; b.t_transpose = t2 - t1; b.t_tree = t3 - t2;
    return rc;
}

/* from_values: ifft each column, then from_coeffs.  values: [C][n] column-major. */
static int orc_batch_from_values(OrcBatch &b, const u64 *values, int cap_height) {
    size_t n = (size_t)1 << b.degree_log;
    double t0 = orc_now();
    b.coeffs.resize((size_t)b.num_polys * n);
    #pragma omp parallel for schedule(dynamic)
    for (int c = 0; c < b.num_polys; c++) {
        u64 *dst = &b.coeffs[(size_t)c * n];
        for (size_t i = 0; i < n; i++) dst[i] = gl_canon(values[(size_t)c * n + i]);
        orc_ifft_inplace(dst, b.degree_log);
    }
    b.t_ifft = orc_now() - t0;
    return orc_batch_from_coeffs(b, cap_height);
}

#endif
