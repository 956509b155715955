/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Poseidon permutation over Goldilocks and the sponge built on it.
 * Restates [DEP plonky2:hash/poseidon.rs, hash/poseidon_goldilocks.rs, hash/hashing.rs].
 * PINNED: the permutation reproduces the 4 upstream known-answer vectors in both the
 * naive form and the fast-partial-round form (tests/test_oracle.py).  The sponge MODE (overwrite, rate 8, no
 * padding) is restated from the published algorithm: parity unpinned.
 */
#ifndef ORACLE_POSEIDON_H
#define ORACLE_POSEIDON_H
#include "gl.h"
#include "poseidon_consts.h"

static const u64 ORC_MDS_CIRC[12] = POSEIDON_MDS_CIRC_INIT;

static inline u64 orc_sbox7(u64 x) {
    u64 x2 = gl_sqr(x), x4 = gl_sqr(x2), x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

/* out[r] = sum_i in[(i+r)%12]*CIRC[i] + in[r]*DIAG[r]   [poseidon.rs::mds_layer] */
static inline void orc_mds(u64 s[12]) {
    /* on 32-bit halves, as plonky2's mds_row_shf does: every sum is < 2^42, no reduction inside */
    u64 lo[24], hi[24], o[12];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = s[i] & GL_EPS; hi[i] = hi[i + 12] = s[i] >> 32; }
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
        for (int i = 0; i < 12; i++) { al += lo[i + r] * ORC_MDS_CIRC[i]; ah += hi[i + r] * ORC_MDS_CIRC[i]; }
        if (r == 0) { al += lo[0] * POSEIDON_MDS_DIAG0; ah += hi[0] * POSEIDON_MDS_DIAG0; }
        o[r] = gl_reduce128((u128)al + ((u128)ah << 32));
    }
    for (int r = 0; r < 12; r++) s[r] = o[r];
}

/* Reference ("naive") form: every round adds 12 constants, S-box (all lanes / lane 0), MDS. */
static inline void orc_poseidon_naive(u64 s[12]) {
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
    int rnd = 0;
    for (int ph = 0; ph < 3; ph++) {
        int n = ph == 1 ? POSEIDON_PARTIAL_ROUNDS : POSEIDON_HALF_FULL_ROUNDS;
        for (int k = 0; k < n; k++, rnd++) {
            for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC[12 * rnd + i]);
            if (ph == 1) s[0] = orc_sbox7(s[0]);
            else for (int i = 0; i < 12; i++) s[i] = orc_sbox7(s[i]);
            orc_mds(s);
        }
    }
}

/* ---- lazily reduced helpers of the production form: values are ANY u64 representative, canonicalised once at the end
 * (plonky2 does the same: GoldilocksField is non-canonical internally).  Only orc_poseidon uses them. ---- */
static inline u64 orc_red128(u128 x) {            /* gl_reduce128 without the final canonicalisation */
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0;
    u64 borrow = __builtin_sub_overflow(lo, hh, &t0);
    t0 -= (0 - borrow) & GL_EPS;
    u64 t1 = hl * GL_EPS;
    u64 t2;
    u64 carry = __builtin_add_overflow(t0, t1, &t2);
    t2 += (0 - carry) & GL_EPS;
    return t2;
}
static inline u64 orc_mulr(u64 a, u64 b) { return orc_red128((u128)a * b); }
static inline u64 orc_sbox7r(u64 x) {
    u64 x2 = orc_mulr(x, x), x4 = orc_mulr(x2, x2), x3 = orc_mulr(x, x2);
    return orc_mulr(x3, x4);
}
static inline u64 orc_addc(u64 s, u64 c) {         /* s any u64, c canonical: one wrap at most */
    u64 t = s + c;
    return t + (GL_EPS & (0 - (u64)(t < s)));      /* branch-free: the carry is a coin flip */
}
/* y = 4 * circ(CIRC) x, out[r] = sum_i x[(i+r)%12] * CIRC[i], as a split correlation: 12 -> cyclic 6 + negacyclic 6, the cyclic
 * half again 3 + 3: 54 products instead of 144.  plonky2 does the same job on x86-64 with its FFT-based `mds_multiply_freq`
 * on the 32-bit halves [DEP plonky2:hash/poseidon_goldilocks.rs]; this is the equivalent, cheaper MDS of the production
 * form.  orc_mds above stays the direct 144-product form of the independent (naive) permutation. */
static inline void orc_circ12_x4(const long long x[12], long long y[12]) {
    static const long long CPP[3] = {
        (long long)(ORC_MDS_CIRC[0] + ORC_MDS_CIRC[3] + ORC_MDS_CIRC[6] + ORC_MDS_CIRC[9]),
        (long long)(ORC_MDS_CIRC[1] + ORC_MDS_CIRC[4] + ORC_MDS_CIRC[7] + ORC_MDS_CIRC[10]),
        (long long)(ORC_MDS_CIRC[2] + ORC_MDS_CIRC[5] + ORC_MDS_CIRC[8] + ORC_MDS_CIRC[11])};
    static const long long CPQ[3] = {
        (long long)ORC_MDS_CIRC[0] - (long long)ORC_MDS_CIRC[3] + (long long)ORC_MDS_CIRC[6] - (long long)ORC_MDS_CIRC[9],
        (long long)ORC_MDS_CIRC[1] - (long long)ORC_MDS_CIRC[4] + (long long)ORC_MDS_CIRC[7] - (long long)ORC_MDS_CIRC[10],
        (long long)ORC_MDS_CIRC[2] - (long long)ORC_MDS_CIRC[5] + (long long)ORC_MDS_CIRC[8] - (long long)ORC_MDS_CIRC[11]};
    static const long long CQ[6] = {
        (long long)ORC_MDS_CIRC[0] - (long long)ORC_MDS_CIRC[6], (long long)ORC_MDS_CIRC[1] - (long long)ORC_MDS_CIRC[7],
        (long long)ORC_MDS_CIRC[2] - (long long)ORC_MDS_CIRC[8], (long long)ORC_MDS_CIRC[3] - (long long)ORC_MDS_CIRC[9],
        (long long)ORC_MDS_CIRC[4] - (long long)ORC_MDS_CIRC[10], (long long)ORC_MDS_CIRC[5] - (long long)ORC_MDS_CIRC[11]};
    long long P[6], Q[6], PP[3], PQ[3], U[6];
    for (int j = 0; j < 6; j++) { P[j] = x[j] + x[j + 6]; Q[j] = x[j] - x[j + 6]; }
    for (int j = 0; j < 3; j++) { PP[j] = P[j] + P[j + 3]; PQ[j] = P[j] - P[j + 3]; }
    for (int r = 0; r < 3; r++) {
        long long uu = 0, uv = 0;
        for (int i = 0; i < 3; i++) uu += PP[(i + r) % 3] * CPP[i];
        for (int i = 0; i < 3; i++) uv += (i + r >= 3 ? -PQ[(i + r) % 3] : PQ[(i + r) % 3]) * CPQ[i];
        U[r] = uu + uv;
        U[r + 3] = uu - uv;
    }
    for (int r = 0; r < 6; r++) {
        long long v = 0;
        for (int i = 0; i < 6; i++) v += (i + r >= 6 ? -Q[(i + r) % 6] : Q[(i + r) % 6]) * CQ[i];
        y[r] = U[r] + 2 * v;
        y[r + 6] = U[r] - 2 * v;
    }
}
static inline void orc_mds_r(u64 s[12]) {
    /* on 32-bit halves: every true sum is in [0, 2^42), the x4 intermediates below 2^45 in magnitude */
    long long lo[12], hi[12], yl[12], yh[12];
    for (int i = 0; i < 12; i++) { lo[i] = (long long)(s[i] & GL_EPS); hi[i] = (long long)(s[i] >> 32); }
    orc_circ12_x4(lo, yl);
    orc_circ12_x4(hi, yh);
    yl[0] += 4 * (long long)POSEIDON_MDS_DIAG0 * lo[0];
    yh[0] += 4 * (long long)POSEIDON_MDS_DIAG0 * hi[0];
    for (int r = 0; r < 12; r++) s[r] = orc_red128((u128)(u64)(yl[r] >> 2) + ((u128)(u64)(yh[r] >> 2) << 32));
}
/* sum of up to 12 products of u64s, accumulated as (low words, high words): no carry tracking, 2^64 = eps at the end
 * (what plonky2's reduce_u160 / mds_partial_layer_fast achieve with a u160 accumulator) */
#define ORC_DOT_ACC(alo, ahi, a, b) do { u128 t__ = (u128)(a) * (b); (alo) += (u64)t__; (ahi) += (u64)(t__ >> 64); } while (0)
static inline u64 orc_dot_finish(u128 alo, u128 ahi) { return orc_red128((u128)orc_red128(alo) + (u128)orc_red128(ahi) * GL_EPS); }

/* Production form of plonky2 (partial rounds through the sparse factorisation); identical outputs.  The timed CPU
 * baseline runs this one; orc_poseidon_naive above is the independent form it is checked against (tests/test_oracle.py). */
static inline void orc_poseidon(u64 s[12]) {
    int rnd = 0;
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = orc_sbox7r(orc_addc(s[i], POSEIDON_RC[12 * rnd + i]));
        orc_mds_r(s);
    }
    for (int i = 0; i < 12; i++) s[i] = orc_addc(s[i], POSEIDON_FAST_FIRST[i]);
    {
        u64 o[11];
        for (int i = 0; i < 11; i++) {
            u128 alo = 0, ahi = 0;
            for (int j = 0; j < 11; j++) ORC_DOT_ACC(alo, ahi, POSEIDON_FAST_INIT[11 * i + j], s[j + 1]);
            o[i] = orc_dot_finish(alo, ahi);
        }
        for (int i = 0; i < 11; i++) s[i + 1] = o[i];
    }
    for (int r = 0; r < POSEIDON_PARTIAL_ROUNDS; r++) {
        const u64 s0 = orc_addc(orc_sbox7r(s[0]), POSEIDON_FAST_K[r]);
        u128 alo = (u128)s0 * 25, ahi = 0;
        for (int i = 0; i < 11; i++) ORC_DOT_ACC(alo, ahi, POSEIDON_FAST_ROW[11 * r + i], s[i + 1]);
        const u64 d = orc_dot_finish(alo, ahi);
        for (int i = 0; i < 11; i++) s[i + 1] = orc_red128((u128)POSEIDON_FAST_COL[11 * r + i] * s0 + s[i + 1]);
        s[0] = d;
    }
    rnd += POSEIDON_PARTIAL_ROUNDS;
    for (int k = 0; k < 4; k++, rnd++) {
        for (int i = 0; i < 12; i++) s[i] = orc_sbox7r(orc_addc(s[i], POSEIDON_RC[12 * rnd + i]));
        orc_mds_r(s);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

/* hash_n_to_m_no_pad with m = 4: overwrite-mode sponge, rate 8 [hashing.rs] */
static inline void orc_hash_no_pad(const u64 *in, size_t n, u64 out[4]) {
    u64 s[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < len; i++) s[i] = gl_canon(in[off + i]);
        orc_poseidon(s);
    }
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

/* hash_or_noop: <= 4 elements are zero-padded into the digest without hashing [hashing.rs / config.rs] */
static inline void orc_hash_or_noop(const u64 *in, size_t n, u64 out[4]) {
    if (n <= 4) {
        for (size_t i = 0; i < 4; i++) out[i] = i < n ? gl_canon(in[i]) : 0;
    } else {
        orc_hash_no_pad(in, n, out);
    }
}

/* two_to_one = permute([l, r, 0,0,0,0])[0..4] [hashing.rs::compress] */
static inline void orc_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    orc_poseidon(s);
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

#endif
