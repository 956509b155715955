/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  PARITY UNPINNED: restated from the published algorithm of
 * plonky2 0.1.4 (source absent here); self-consistency is checked by proving and then VERIFYING with a restatement of
 * plonky2's own verifier, as the reference's tests do (/root/reference/eth-lc-plonky2/src/unit_tests.rs:29-35).
 *
 * Batched FRI over F_p^2 and the opening set:
 *   [DEP plonky2:plonk/proof.rs::OpeningSet::new]                       evaluate every polynomial at zeta / g*zeta
 *   [DEP plonky2:fri/oracle.rs::PolynomialBatch::prove_openings]         sum alpha^j f_j, divide_by_linear, combine, LDE
 *   [DEP plonky2:fri/prover.rs::{fri_committed_trees, fri_proof_of_work, fri_prover_query_rounds}]
 *   [DEP plonky2:fri/verifier.rs::{verify_fri_proof, fri_combine_initial, compute_evaluation}]
 *   [DEP plonky2:fri/reduction_strategies.rs::ConstantArityBits]        
 * The prover here follows plonky2's COEFFICIENT-space route (Horner scan for divide_by_linear, coset FFT per fold
 * round); the CUDA engine computes the same objects in the evaluation domain, so agreement is a real cross-check.
 * SECOND RESTATEMENT: tests/golden/plonk_restatement.py (pure Python: synthetic division in coefficient space, evaluation
 * by Horner at every coset point instead of FFTs, Lagrange interpolation in the verifier's fold) reproduces this file's
 * FriProof word for word, one reduction round included (tests/golden/plonk_proof.json).
 * Proof-of-work: plonky2 accepts any witness (rayon find_any); parity is defined on the SMALLEST valid witness.
 */
#ifndef ORACLE_FRI_H
#define ORACLE_FRI_H
#include <vector>
#include "challenger.h"
#include "commit.h"

typedef std::vector<gl2> vec2;

struct OrcFriParams {
    int degree_bits = 0, rate_bits = 3, cap_height = 4, pow_bits = 16, num_query_rounds = 28;
    std::vector<int> arity_bits;
};
/* ConstantArityBits(arity_bits, final_poly_bits) */
static inline std::vector<int> orc_fri_reduction_arity_bits(int degree_bits, int rate_bits, int cap_height, int arity_bits,
                                                            int final_poly_bits) {
    std::vector<int> r;
    while (degree_bits > final_poly_bits && degree_bits + rate_bits - arity_bits >= cap_height) {
        r.push_back(arity_bits);
        degree_bits -= arity_bits;
    }
    return r;
}

struct OrcFriPolyRef { int oracle, poly; };
struct OrcFriBatchInfo { gl2 point; std::vector<OrcFriPolyRef> polys; };

struct OrcFriQueryStep { vec2 evals; vec64 path; };
struct OrcFriQueryRound {
    std::vector<vec64> leaves, paths; /* per initial oracle */
    std::vector<OrcFriQueryStep> steps;
};
struct OrcFriProof {
    std::vector<vec64> caps; /* commit_phase_merkle_caps */
    std::vector<OrcFriQueryRound> rounds;
    vec2 final_poly;
    u64 pow_witness = 0;
};

/* p.to_extension().eval(z): Horner over F_p^2 */
static inline gl2 orc_eval_base_poly(const u64 *c, size_t n, gl2 z) {
    gl2 acc = gl2_from(0);
    for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, z), gl2_from(c[i]));
    return acc;
}
static inline gl2 orc_eval_ext_poly(const vec2 &c, gl2 z) {
    gl2 acc = gl2_from(0);
    for (size_t i = c.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, z), c[i]);
    return acc;
}
/* eval_commitment(z, batch): every polynomial of the batch at z */
static inline vec2 orc_batch_eval(const OrcBatch &b, gl2 z) {
    size_t n = (size_t)1 << b.degree_log;
    vec2 out(b.num_polys);
    #pragma omp parallel for schedule(dynamic)
    for (int c = 0; c < b.num_polys; c++) out[c] = orc_eval_base_poly(&b.coeffs[(size_t)c * n], n, z);
    return out;
}

/* fft over F_p^2 with base-field roots: component-wise */
static inline void orc_coset_fft_ext(vec2 &v, int log_n, u64 shift) {
    size_t n = (size_t)1 << log_n;
    vec64 a(n), b(n);
    u64 pw = 1;
    for (size_t k = 0; k < n; k++) { a[k] = gl_mul(v[k].a, pw); b[k] = gl_mul(v[k].b, pw); pw = gl_mul(pw, shift); }
    orc_fft_inplace(a.data(), log_n);
    orc_fft_inplace(b.data(), log_n);
    for (size_t k = 0; k < n; k++) v[k] = gl2_make(a[k], b[k]);
}

static inline void orc_observe_cap(OrcChallenger *ch, const vec64 &cap) { orc_ch_observe_n(ch, cap.data(), cap.size()); }

/* smallest w such that the duplex of (buffered inputs, w) has >= pow_bits leading zeros in state[7] */
static inline u64 orc_fri_pow(OrcChallenger *ch, int pow_bits) {
    u64 base[12];
    memcpy(base, ch->state, sizeof(base));
    for (int i = 0; i < ch->n_in; i++) base[i] = ch->in_buf[i];
    int pos = ch->n_in;
    u64 found = ~0ULL;
    const u64 chunk = 1 << 14;
    for (u64 start = 0; found == ~0ULL; start += chunk) {
        u64 best = ~0ULL;
        #pragma omp parallel for reduction(min : best)
        for (u64 w = start; w < start + chunk; w++) {
            u64 s[12];
            memcpy(s, base, sizeof(s));
            s[pos] = w;
            orc_poseidon(s);
            if (__builtin_clzll(s[7] | 1) >= pow_bits && (s[7] >> (64 - pow_bits)) == 0) { if (w < best) best = w; }
        }
        found = best;
    }
    orc_ch_observe(ch, found);
    u64 resp = orc_ch_challenge(ch);
    assert((resp >> (64 - pow_bits)) == 0);
    return found;
}

/* PolynomialBatch::prove_openings */
static OrcFriProof orc_fri_prove(const std::vector<OrcFriBatchInfo> &batches, const std::vector<const OrcBatch *> &oracles,
                                 OrcChallenger *ch, const OrcFriParams &P) {
    const int log_n = P.degree_bits, log_l = log_n + P.rate_bits;
    const size_t n = (size_t)1 << log_n, L = (size_t)1 << log_l;
    gl2 alpha = orc_ch_challenge_ext(ch);
    vec2 final_poly(n, gl2_from(0));
    for (const OrcFriBatchInfo &bt : batches) {
        /* composition = sum_j alpha^j f_j  (reduce_polys_base) */
        vec2 comp(n, gl2_from(0));
        gl2 apow = gl2_from(1);
        for (const OrcFriPolyRef &pr : bt.polys) {
            const u64 *c = &oracles[pr.oracle]->coeffs[(size_t)pr.poly * n];
            #pragma omp parallel for schedule(static)
            for (size_t k = 0; k < n; k++) comp[k] = gl2_add(comp[k], gl2_scale(apow, c[k]));
            apow = gl2_mul(apow, alpha);
        }
        /* divide_by_linear(point): Horner scan from the top, drop the remainder, pad with a zero */
        vec2 quot(n, gl2_from(0));
        gl2 acc = gl2_from(0);
        for (size_t k = n; k-- > 0;) {
            acc = gl2_add(gl2_mul(acc, bt.point), comp[k]);
            if (k > 0) quot[k - 1] = acc;
        }
        /* alpha.shift_poly(final) : final *= alpha^count (count = this batch's size); final += quotient */
        gl2 sh = gl2_pow(alpha, bt.polys.size());
        #pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) final_poly[k] = gl2_add(gl2_mul(final_poly[k], sh), quot[k]);
    }
    /* lde + coset_fft(7) */
    vec2 coeffs(L, gl2_from(0));
    for (size_t k = 0; k < n; k++) coeffs[k] = final_poly[k];
    vec2 values = coeffs;
    orc_coset_fft_ext(values, log_l, GL_GENERATOR);

    OrcFriProof proof;
    std::vector<OrcMerkleTree> trees;
    u64 shift = GL_GENERATOR;
    int cur_log = log_l;
    for (int arity_bits : P.arity_bits) {
        size_t len = (size_t)1 << cur_log, arity = (size_t)1 << arity_bits;
        OrcMerkleTree t;
        t.num_leaves = len >> arity_bits; t.leaf_len = 2 * arity; t.cap_height = P.cap_height;
        t.leaves.resize(2 * len);
        for (size_t k = 0; k < len; k++) {  /* reverse_index_bits, then chunks of `arity`, flattened */
            gl2 v = values[bitrev(k, cur_log)];
            t.leaves[2 * k] = v.a; t.leaves[2 * k + 1] = v.b;
        }
        int rc = orc_merkle_build(t);
        assert(rc == 0); (void)rc;
        orc_observe_cap(ch, t.cap);
        proof.caps.push_back(t.cap);
        trees.push_back(std::move(t));
        gl2 beta = orc_ch_challenge_ext(ch);
        vec2 next(len >> arity_bits);
        for (size_t k = 0; k < next.size(); k++) {  /* reduce_with_powers(chunk, beta) */
            gl2 a = gl2_from(0);
            for (size_t t2 = arity; t2-- > 0;) a = gl2_add(gl2_mul(a, beta), coeffs[k * arity + t2]);
            next[k] = a;
        }
        coeffs.swap(next);
        shift = gl_pow(shift, arity);
        cur_log -= arity_bits;
        values = coeffs;
        orc_coset_fft_ext(values, cur_log, shift);
    }
    coeffs.resize(coeffs.size() >> P.rate_bits);
    for (const gl2 &c : coeffs) orc_ch_observe_ext(ch, c);
    proof.final_poly = coeffs;
    proof.pow_witness = orc_fri_pow(ch, P.pow_bits);
    for (int q = 0; q < P.num_query_rounds; q++) {
        size_t x_index = orc_ch_challenge(ch) % L;
        OrcFriQueryRound qr;
        for (const OrcBatch *o : oracles) {
            const OrcMerkleTree &t = o->tree;
            qr.leaves.emplace_back(t.leaves.begin() + x_index * t.leaf_len, t.leaves.begin() + (x_index + 1) * t.leaf_len);
            vec64 path(4 * 64);
            int k = orc_merkle_prove(t, x_index, path.data());
            path.resize(4 * k);
            qr.paths.push_back(path);
        }
        size_t xi = x_index;
        for (size_t i = 0; i < trees.size(); i++) {
            int ab = P.arity_bits[i];
            const OrcMerkleTree &t = trees[i];
            size_t c = xi >> ab;
            OrcFriQueryStep st;
            for (size_t e = 0; e < ((size_t)1 << ab); e++) st.evals.push_back(gl2_make(t.leaves[c * t.leaf_len + 2 * e], t.leaves[c * t.leaf_len + 2 * e + 1]));
            st.path.resize(4 * 64);
            int k = orc_merkle_prove(t, c, st.path.data());
            st.path.resize(4 * k);
            qr.steps.push_back(st);
            xi = c;
        }
        proof.rounds.push_back(qr);
    }
    return proof;
}

/* compute_evaluation: interpolate {(x g^i, P(x g^i))} over the coset of x and evaluate at beta */
static inline gl2 orc_fri_compute_evaluation(u64 x, size_t x_index_within_coset, int arity_bits, const vec2 &evals_in, gl2 beta) {
    size_t arity = (size_t)1 << arity_bits;
    u64 g = gl_root_of_unity(arity_bits);
    vec2 evals(arity);
    for (size_t i = 0; i < arity; i++) evals[bitrev(i, arity_bits)] = evals_in[i];
    size_t rev = bitrev(x_index_within_coset, arity_bits);
    u64 coset_start = gl_mul(x, gl_pow(g, arity - rev));
    /* Lagrange interpolation at beta over points coset_start * g^i */
    std::vector<u64> pts(arity);
    u64 y = 1;
    for (size_t i = 0; i < arity; i++) { pts[i] = gl_mul(coset_start, y); y = gl_mul(y, g); }
    gl2 sum = gl2_from(0);
    for (size_t i = 0; i < arity; i++) {
        gl2 num = gl2_from(1);
        u64 den = 1;
        for (size_t j = 0; j < arity; j++) {
            if (j == i) continue;
            num = gl2_mul(num, gl2_sub(beta, gl2_from(pts[j])));
            den = gl_mul(den, gl_sub(pts[i], pts[j]));
        }
        sum = gl2_add(sum, gl2_mul(evals[i], gl2_scale(num, gl_inv(den))));
    }
    return sum;
}

/* verify_fri_proof.  `openings[b]` = claimed values of batch b's polynomials at its point, in instance order.
 * `ch` must be in the state right after observing the openings.  Returns 0 when the proof verifies, else a
 * positive code naming the failed check. */
static int orc_fri_verify(const std::vector<OrcFriBatchInfo> &batches, const std::vector<vec2> &openings,
                          const std::vector<vec64> &initial_caps, const std::vector<int> &oracle_leaf_len, const OrcFriProof &proof,
                          OrcChallenger *ch, const OrcFriParams &P) {
    const int log_l = P.degree_bits + P.rate_bits;
    const size_t L = (size_t)1 << log_l;
    if (proof.caps.size() != P.arity_bits.size() || (int)proof.rounds.size() != P.num_query_rounds) return 1;
    gl2 alpha = orc_ch_challenge_ext(ch);
    std::vector<gl2> betas;
    for (const vec64 &cap : proof.caps) { orc_observe_cap(ch, cap); betas.push_back(orc_ch_challenge_ext(ch)); }
    for (const gl2 &c : proof.final_poly) orc_ch_observe_ext(ch, c);
    orc_ch_observe(ch, proof.pow_witness);
    u64 pow_resp = orc_ch_challenge(ch);
    if ((pow_resp >> (64 - P.pow_bits)) != 0) return 2;
    /* PrecomputedReducedOpenings */
    std::vector<gl2> reduced;
    for (const vec2 &vals : openings) {
        gl2 acc = gl2_from(0);
        for (size_t i = vals.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, alpha), vals[i]);
        reduced.push_back(acc);
    }
    for (const OrcFriQueryRound &qr : proof.rounds) {
        size_t x_index = orc_ch_challenge(ch) % L;
        if (qr.leaves.size() != initial_caps.size()) return 3;
        for (size_t o = 0; o < initial_caps.size(); o++) {
            if ((int)qr.leaves[o].size() != oracle_leaf_len[o]) return 3;
            if (!orc_merkle_verify(qr.leaves[o].data(), qr.leaves[o].size(), x_index, initial_caps[o].data(), qr.paths[o].data(), (int)qr.paths[o].size() / 4)) return 4;
        }
        u64 subgroup_x = gl_mul(GL_GENERATOR, gl_pow(gl_root_of_unity(log_l), bitrev(x_index, log_l)));
        /* fri_combine_initial */
        gl2 sum = gl2_from(0);
        for (size_t b = 0; b < batches.size(); b++) {
            gl2 acc = gl2_from(0);
            const auto &polys = batches[b].polys;
            for (size_t i = polys.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, alpha), gl2_from(qr.leaves[polys[i].oracle][polys[i].poly]));
            gl2 num = gl2_sub(acc, reduced[b]);
            gl2 den = gl2_sub(gl2_from(subgroup_x), batches[b].point);
            sum = gl2_add(gl2_mul(sum, gl2_pow(alpha, polys.size())), gl2_mul(num, gl2_inv(den)));
        }
        gl2 old_eval = sum;
        size_t xi = x_index;
        for (size_t i = 0; i < P.arity_bits.size(); i++) {
            int ab = P.arity_bits[i];
            size_t arity = (size_t)1 << ab, coset_index = xi >> ab, within = xi & (arity - 1);
            const OrcFriQueryStep &st = qr.steps[i];
            if (st.evals.size() != arity) return 3;
            if (!gl2_eq(st.evals[within], old_eval)) return 5;
            old_eval = orc_fri_compute_evaluation(subgroup_x, within, ab, st.evals, betas[i]);
            vec64 flat(2 * arity);
            for (size_t e = 0; e < arity; e++) { flat[2 * e] = st.evals[e].a; flat[2 * e + 1] = st.evals[e].b; }
            if (!orc_merkle_verify(flat.data(), flat.size(), coset_index, proof.caps[i].data(), st.path.data(), (int)st.path.size() / 4)) return 6;
            for (int k = 0; k < ab; k++) subgroup_x = gl_sqr(subgroup_x);
            xi = coset_index;
        }
        if (!gl2_eq(orc_eval_ext_poly(proof.final_poly, gl2_from(subgroup_x)), old_eval)) return 7;
    }
    return 0;
}

/* Flat u64 serialisation shared by the oracle and the engine (NOT plonky2's wire format -- that is row f4):
 * [R] R x cap(2^h * 4) | [F] F x (a, b) | pow_witness | [Q] Q x { [O] O x { [len] leaf | [plen] path } | R x { 2*arity evals | [plen] path } } */
static inline vec64 orc_fri_proof_blob(const OrcFriProof &p) {
    vec64 o;
    o.push_back(p.caps.size());
    for (const vec64 &c : p.caps) { o.push_back(c.size()); o.insert(o.end(), c.begin(), c.end()); }
    o.push_back(p.final_poly.size());
    for (const gl2 &c : p.final_poly) { o.push_back(c.a); o.push_back(c.b); }
    o.push_back(p.pow_witness);
    o.push_back(p.rounds.size());
    for (const OrcFriQueryRound &q : p.rounds) {
        o.push_back(q.leaves.size());
        for (size_t i = 0; i < q.leaves.size(); i++) {
            o.push_back(q.leaves[i].size()); o.insert(o.end(), q.leaves[i].begin(), q.leaves[i].end());
            o.push_back(q.paths[i].size() / 4); o.insert(o.end(), q.paths[i].begin(), q.paths[i].end());
        }
        o.push_back(q.steps.size());
        for (const OrcFriQueryStep &s : q.steps) {
            o.push_back(s.evals.size());
            for (const gl2 &e : s.evals) { o.push_back(e.a); o.push_back(e.b); }
            o.push_back(s.path.size() / 4); o.insert(o.end(), s.path.begin(), s.path.end());
        }
    }
    return o;
}
static inline bool orc_fri_proof_from_blob(const u64 *b, size_t len, OrcFriProof &p) {
    size_t i = 0;
    auto take = [&](size_t k) { size_t r = i; i += k; return r <= len && i <= len ? r : (size_t)-1; };
    #define RD(var) { size_t q_ = take(1); if (q_ == (size_t)-1) return false; var = b[q_]; }
    u64 R, F, Q, O, S, n;
    RD(R);
    for (u64 r = 0; r < R; r++) { RD(n); size_t q = take(n); if (q == (size_t)-1) return false; p.caps.emplace_back(b + q, b + q + n); }
    RD(F);
    for (u64 f = 0; f < F; f++) { u64 a, c; RD(a); RD(c); p.final_poly.push_back(gl2_make(a, c)); }
    RD(p.pow_witness);
    RD(Q);
    for (u64 q = 0; q < Q; q++) {
        OrcFriQueryRound qr;
        RD(O);
        for (u64 o = 0; o < O; o++) {
            RD(n); size_t s = take(n); if (s == (size_t)-1) return false; qr.leaves.emplace_back(b + s, b + s + n);
            RD(n); s = take(4 * n); if (s == (size_t)-1) return false; qr.paths.emplace_back(b + s, b + s + 4 * n);
        }
        RD(S);
        for (u64 s2 = 0; s2 < S; s2++) {
            OrcFriQueryStep st;
            RD(n);
            for (u64 e = 0; e < n; e++) { u64 a, c; RD(a); RD(c); st.evals.push_back(gl2_make(a, c)); }
            RD(n); size_t s = take(4 * n); if (s == (size_t)-1) return false; st.path.assign(b + s, b + s + 4 * n);
            qr.steps.push_back(st);
        }
        p.rounds.push_back(qr);
    }
    #undef RD
    return i == len;
}
#endif
