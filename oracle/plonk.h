/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  PARITY UNPINNED (tier C of SURVEY.md Appendix A: structure
 * restated, order-sensitive details cannot be checked against the absent source).  Self-consistency is established the
 * way the reference's tests do it (/root/reference/eth-lc-plonky2/src/unit_tests.rs:29-35): prove, then VERIFY with a
 * restatement of plonky2's verifier that re-evaluates every constraint at zeta over F_p^2.
 * SECOND RESTATEMENT: tests/golden/plonk_restatement.py (pure Python, written from the published algorithm with different
 * methods -- direct DFT sums, naive PoseidonGate rounds, level-by-level Merkle trees) proves the same circuits; its proofs
 * (tests/golden/plonk_proof.json) equal this file's word for word, and its verifier accepts this file's proofs
 * (tests/test_plonk_cpu.py::test_oracle_proof_equals_python_restatement).  Still not plonky2 output: the cap stays.
 *
 *   [DEP plonky2:plonk/prover.rs::{prove_with_partition_witness, wires_permutation_partial_products_and_zs,
 *        compute_quotient_polys}]                                            
 *   [DEP plonky2:plonk/vanishing_poly.rs::{eval_vanishing_poly(_base_batch), evaluate_gate_constraints(_base_batch)}]
 *   [DEP plonky2:plonk/plonk_common.rs::{check_partial_products, ZeroPolyOnCoset, reduce_with_powers_multi}]
 *   [DEP plonky2:gates/{noop,constant,public_input,arithmetic_base,poseidon}.rs, gates/selectors.rs]
 *   [DEP plonky2:plonk/verifier.rs::verify_with_challenges, plonk/get_challenges.rs]
 * Gate set: the five core gates restated in SURVEY.md A.8 (Noop, Constant, PublicInput, Arithmetic, Poseidon) --
 * enough for a circuit-shaped synthetic proof; the other gates of the real eth-lc circuit need their source (Appendix D).
 */
#ifndef ORACLE_PLONK_H
#define ORACLE_PLONK_H
#include "fri.h"
#include "gates.h"

enum { ORC_G_NOOP = 0, ORC_G_CONSTANT = 1, ORC_G_PUBLIC_INPUT = 2, ORC_G_ARITHMETIC = 3, ORC_G_POSEIDON = 4 };
#define ORC_UNUSED_SELECTOR 0xFFFFFFFFULL

struct OrcGateInfo { int kind, selector_index, group_start, group_end; int p[4]; };
struct OpsBase;
template <class O, class Emit> void orc_eval_gate(int kind, const int p[4], const typename O::T *w, const typename O::T *c, const u64 pi_hash[4], Emit emit);

struct OrcCircuit {
    int degree_bits = 0, num_wires = 135, num_routed = 80, num_gate_constants = 2, num_selectors = 0;
    int num_challenges = 2, quotient_degree_factor = 8, rate_bits = 3, cap_height = 4, pow_bits = 16, num_query_rounds = 28;
    std::vector<OrcGateInfo> gates;   /* sorted by degree as plonky2 does */
    vec64 k_is;                       /* 7^j */
    u64 circuit_digest[4] = {0, 0, 0, 0};
    int num_constants() const { return num_selectors + num_gate_constants; }
    int num_partial_products() const { return (num_routed + quotient_degree_factor - 1) / quotient_degree_factor - 1; }
    int num_zs_pp() const { return num_challenges * (1 + num_partial_products()); }
    int num_quotient_polys() const { return num_challenges * quotient_degree_factor; }
    int num_gate_constraints() const;   /* max over the gates (counted by running each evaluator once) */
};

/* ---- field-generic helpers: T = u64 (base, canonical) or gl2 ---- */
struct OpsBase {
    typedef u64 T;
    static T from(u64 x) { return x; }
    static T add(T a, T b) { return gl_add(a, b); }
    static T sub(T a, T b) { return gl_sub(a, b); }
    static T mul(T a, T b) { return gl_mul(a, b); }
    static T scale(T a, u64 s) { return gl_mul(a, s); }
};
struct OpsExt {
    typedef gl2 T;
    static T from(u64 x) { return gl2_from(x); }
    static T add(T a, T b) { return gl2_add(a, b); }
    static T sub(T a, T b) { return gl2_sub(a, b); }
    static T mul(T a, T b) { return gl2_mul(a, b); }
    static T scale(T a, u64 s) { return gl2_scale(a, s); }
};

/* compute_filter(row, group, s, many_selectors) = prod_{i in group, i != row} (i - s) * (UNUSED - s if many) */
template <class O> typename O::T orc_compute_filter(int row, int gs, int ge, typename O::T s, bool many) {
    typename O::T f = O::from(1);
    for (int i = gs; i < ge; i++) if (i != row) f = O::mul(f, O::sub(O::from((u64)i), s));
    if (many) f = O::mul(f, O::sub(O::from(ORC_UNUSED_SELECTOR), s));
    return f;
}

/* Poseidon round pieces over a generic field (constant_layer_field, sbox, mds_layer_field, fast partial layers) */
template <class O> void orc_psd_mds(typename O::T s[12]) {
    typename O::T o[12];
    for (int r = 0; r < 12; r++) {
        typename O::T acc = O::from(0);
        for (int i = 0; i < 12; i++) acc = O::add(acc, O::scale(s[(i + r) % 12], ORC_MDS_CIRC[i]));
        if (r == 0) acc = O::add(acc, O::scale(s[0], POSEIDON_MDS_DIAG0));
        o[r] = acc;
    }
    for (int r = 0; r < 12; r++) s[r] = o[r];
}
template <class O> typename O::T orc_psd_sbox(typename O::T x) {
    typename O::T x2 = O::mul(x, x), x4 = O::mul(x2, x2), x3 = O::mul(x, x2);
    return O::mul(x3, x4);
}

/* Gate::eval_unfiltered.  `w` = local wires, `c` = local constants AFTER the selector prefix, emit(t) receives the
 * constraints in order. */
template <class O, class Emit>
void orc_eval_gate(int kind, const int p[4], const typename O::T *w, const typename O::T *c, const u64 pi_hash[4], Emit emit) {
    typedef typename O::T T;
    if (kind == ORC_G_NOOP) return;
    if (kind > ORC_G_POSEIDON) { orc_eval_gate_ext<O>(kind, p, w, c, emit); return; }
    if (kind == ORC_G_CONSTANT) { for (int i = 0; i < p[0]; i++) emit(O::sub(c[i], w[i])); return; }
    if (kind == ORC_G_PUBLIC_INPUT) { for (int i = 0; i < 4; i++) emit(O::sub(w[i], O::from(pi_hash[i]))); return; }
    if (kind == ORC_G_ARITHMETIC) {
        for (int i = 0; i < p[0]; i++) {
            T m0 = w[4 * i], m1 = w[4 * i + 1], ad = w[4 * i + 2], out = w[4 * i + 3];
            T computed = O::add(O::mul(O::mul(m0, m1), c[0]), O::mul(ad, c[1]));
            emit(O::sub(out, computed));
        }
        return;
    }
    /* PoseidonGate: in 0..12, out 12..24, swap 24, delta 25..29, full_sbox_0(r,i) = 29+12(r-1)+i, partial_sbox(r) = 65+r,
     * full_sbox_1(r,i) = 87+12r+i */
    T swap = w[24];
    emit(O::mul(swap, O::sub(swap, O::from(1))));
    for (int i = 0; i < 4; i++) emit(O::sub(O::mul(swap, O::sub(w[i + 4], w[i])), w[25 + i]));
    T st[12];
    for (int i = 0; i < 4; i++) { st[i] = O::add(w[i], w[25 + i]); st[i + 4] = O::sub(w[i + 4], w[25 + i]); }
    for (int i = 8; i < 12; i++) st[i] = w[i];
    int rnd = 0;
    for (int r = 0; r < 4; r++, rnd++) {
        for (int i = 0; i < 12; i++) st[i] = O::add(st[i], O::from(POSEIDON_RC[12 * rnd + i]));
        if (r != 0)
            for (int i = 0; i < 12; i++) { T in = w[29 + 12 * (r - 1) + i]; emit(O::sub(st[i], in)); st[i] = in; }
        for (int i = 0; i < 12; i++) st[i] = orc_psd_sbox<O>(st[i]);
        orc_psd_mds<O>(st);
    }
    for (int i = 0; i < 12; i++) st[i] = O::add(st[i], O::from(POSEIDON_FAST_FIRST[i]));
    {
        T o[11];
        for (int i = 0; i < 11; i++) {
            T acc = O::from(0);
            for (int j = 0; j < 11; j++) acc = O::add(acc, O::scale(st[j + 1], POSEIDON_FAST_INIT[11 * i + j]));
            o[i] = acc;
        }
        for (int i = 0; i < 11; i++) st[i + 1] = o[i];
    }
    for (int r = 0; r < 22; r++) {
        T in = w[65 + r];
        emit(O::sub(st[0], in));
        T s0 = O::add(orc_psd_sbox<O>(in), O::from(POSEIDON_FAST_K[r]));
        T d = O::scale(s0, 25);
        for (int i = 0; i < 11; i++) d = O::add(d, O::scale(st[i + 1], POSEIDON_FAST_ROW[11 * r + i]));
        for (int i = 0; i < 11; i++) st[i + 1] = O::add(st[i + 1], O::scale(s0, POSEIDON_FAST_COL[11 * r + i]));
        st[0] = d;
    }
    rnd += 22;
    for (int r = 0; r < 4; r++, rnd++) {
        for (int i = 0; i < 12; i++) st[i] = O::add(st[i], O::from(POSEIDON_RC[12 * rnd + i]));
        for (int i = 0; i < 12; i++) { T in = w[87 + 12 * r + i]; emit(O::sub(st[i], in)); st[i] = in; }
        for (int i = 0; i < 12; i++) st[i] = orc_psd_sbox<O>(st[i]);
        orc_psd_mds<O>(st);
    }
    for (int i = 0; i < 12; i++) emit(O::sub(st[i], w[12 + i]));
}

inline int OrcCircuit::num_gate_constraints() const {
    int m = 0;
    std::vector<u64> zeros(num_wires + 8, 0);
    const u64 pi[4] = {0, 0, 0, 0};
    for (const OrcGateInfo &g : gates) {
        int k = 0;
        orc_eval_gate<OpsBase>(g.kind, g.p, zeros.data(), zeros.data(), pi, [&](u64) { k++; });
        if (k > m) m = k;
    }
    return m;
}

/* eval_vanishing_poly at one point for all challenges: terms = [L0(x)(Z-1)] ++ [partial-product checks] ++ [gate
 * constraints], reduced with powers of alpha.  consts = all local constants (selectors first), x = evaluation point
 * (the SHIFTED point in the prover).  z_h_x = Z_H(x), l0 = L_0(x). */
template <class O>
void orc_eval_vanishing(const OrcCircuit &C, typename O::T x, typename O::T l0, const typename O::T *consts, const typename O::T *sig,
                        const typename O::T *w, const typename O::T *zs, const typename O::T *zs_next, const typename O::T *pps,
                        const u64 pi_hash[4], const u64 *betas, const u64 *gammas, const u64 *alphas, typename O::T *out) {
    typedef typename O::T T;
    const int npp = C.num_partial_products(), qd = C.quotient_degree_factor, ngc = C.num_gate_constraints();
    std::vector<T> terms;
    for (int ch = 0; ch < C.num_challenges; ch++) terms.push_back(O::mul(l0, O::sub(zs[ch], O::from(1))));
    for (int ch = 0; ch < C.num_challenges; ch++) {
        int nchunks = npp + 1;
        for (int t = 0; t < nchunks; t++) {
            T prev = t == 0 ? zs[ch] : pps[ch * npp + t - 1];
            T next = t == nchunks - 1 ? zs_next[ch] : pps[ch * npp + t];
            T num = O::from(1), den = O::from(1);
            for (int j = t * qd; j < (t + 1) * qd && j < C.num_routed; j++) {
                T s_id = O::scale(x, C.k_is[j]);
                num = O::mul(num, O::add(O::add(w[j], O::scale(s_id, betas[ch])), O::from(gammas[ch])));
                den = O::mul(den, O::add(O::add(w[j], O::scale(sig[j], betas[ch])), O::from(gammas[ch])));
            }
            terms.push_back(O::sub(O::mul(prev, num), O::mul(next, den)));
        }
    }
    std::vector<T> gate_terms(ngc, O::from(0));
    for (size_t gi = 0; gi < C.gates.size(); gi++) {
        const OrcGateInfo &g = C.gates[gi];
        T filter = orc_compute_filter<O>((int)gi, g.group_start, g.group_end, consts[g.selector_index], C.num_selectors > 1);
        int k = 0;
        orc_eval_gate<O>(g.kind, g.p, w, consts + C.num_selectors, pi_hash, [&](T v) { gate_terms[k] = O::add(gate_terms[k], O::mul(filter, v)); k++; });
    }
    for (const T &t : gate_terms) terms.push_back(t);
    for (int ch = 0; ch < C.num_challenges; ch++) {   /* reduce_with_powers: term t gets alpha^t */
        T acc = O::from(0);
        for (size_t t = terms.size(); t-- > 0;) acc = O::add(O::scale(acc, alphas[ch]), terms[t]);
        out[ch] = acc;
    }
}

/* wires_permutation_partial_products_and_zs for all challenges; output in the committed order
 * [Z_0, Z_1, pp_0[0..npp), pp_1[0..npp)] as [num_zs_pp][n].  wires: [num_wires][n]; sigmas: [num_routed][n] values. */
static inline vec64 orc_partial_products(const OrcCircuit &C, const u64 *wires, const u64 *sigmas, const u64 *betas, const u64 *gammas) {
    const size_t n = (size_t)1 << C.degree_bits;
    const int npp = C.num_partial_products(), qd = C.quotient_degree_factor, nch = C.num_challenges, nchunks = npp + 1;
    vec64 out((size_t)C.num_zs_pp() * n);
    std::vector<u64> sub(n);
    u64 g = gl_root_of_unity(C.degree_bits);
    sub[0] = 1;
    for (size_t i = 1; i < n; i++) sub[i] = gl_mul(sub[i - 1], g);
    for (int ch = 0; ch < nch; ch++) {
        vec64 chunk((size_t)nchunks * n);
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            for (int t = 0; t < nchunks; t++) {
                u64 num = 1, den = 1;
                for (int j = t * qd; j < (t + 1) * qd && j < C.num_routed; j++) {
                    u64 wv = gl_canon(wires[(size_t)j * n + i]);
                    num = gl_mul(num, gl_add(gl_add(wv, gl_mul(betas[ch], gl_mul(C.k_is[j], sub[i]))), gammas[ch]));
                    den = gl_mul(den, gl_add(gl_add(wv, gl_mul(betas[ch], gl_canon(sigmas[(size_t)j * n + i]))), gammas[ch]));
                }
                chunk[(size_t)t * n + i] = gl_mul(num, gl_inv(den));   /* product of the chunk's quotients */
            }
        }
        u64 z = 1;
        for (size_t i = 0; i < n; i++) {   /* sequential accumulation over rows */
            out[(size_t)ch * n + i] = z;
            u64 acc = z;
            for (int t = 0; t < nchunks; t++) {
                acc = gl_mul(acc, chunk[(size_t)t * n + i]);
                if (t < npp) out[((size_t)nch + (size_t)ch * npp + t) * n + i] = acc;
            }
            z = acc;
        }
    }
    return out;
}

/* compute_quotient_polys: [num_challenges][8n] coefficients (coset_ifft of the quotient values on 7*H_{8n}).
 * Returns an empty vector if quotient_degree_bits != rate_bits (the only case restated: step = 1). */
static inline vec64 orc_quotient_polys(const OrcCircuit &C, const OrcBatch &cs, const OrcBatch &wires, const OrcBatch &zs_pp,
                                       const u64 pi_hash[4], const u64 *betas, const u64 *gammas, const u64 *alphas) {
    int qdb = 0;
    while ((1 << qdb) < C.quotient_degree_factor) qdb++;
    if (qdb != C.rate_bits) return vec64();
    const int log_l = C.degree_bits + qdb, nch = C.num_challenges, npp = C.num_partial_products();
    const size_t L = (size_t)1 << log_l, n = (size_t)1 << C.degree_bits, next_step = (size_t)1 << qdb;
    /* ZeroPolyOnCoset */
    u64 g_pow_n = gl_pow(GL_GENERATOR, n);
    std::vector<u64> zh(next_step), zh_inv(next_step);
    for (size_t i = 0; i < next_step; i++) { zh[i] = gl_sub(gl_mul(g_pow_n, gl_pow(gl_root_of_unity(qdb), i)), 1); zh_inv[i] = gl_inv(zh[i]); }
    u64 wl = gl_root_of_unity(log_l);
    vec64 vals((size_t)nch * L);
    const int nc = C.num_constants();
    #pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < L; i++) {
        u64 x = gl_mul(GL_GENERATOR, gl_pow(wl, i));
        size_t i_next = (i + next_step) % L;
        const u64 *row_cs = &cs.tree.leaves[bitrev(i, log_l) * cs.tree.leaf_len];
        const u64 *row_w = &wires.tree.leaves[bitrev(i, log_l) * wires.tree.leaf_len];
        const u64 *row_z = &zs_pp.tree.leaves[bitrev(i, log_l) * zs_pp.tree.leaf_len];
        const u64 *row_zn = &zs_pp.tree.leaves[bitrev(i_next, log_l) * zs_pp.tree.leaf_len];
        u64 l0 = gl_mul(zh[i % next_step], gl_inv(gl_mul((u64)n % GL_P, gl_sub(x, 1))));
        u64 res[8];
        orc_eval_vanishing<OpsBase>(C, x, l0, row_cs, row_cs + nc, row_w, row_z, row_zn, row_z + nch, pi_hash, betas, gammas, alphas, res);
        for (int ch = 0; ch < nch; ch++) vals[(size_t)ch * L + i] = gl_mul(res[ch], zh_inv[i % next_step]);
        (void)npp;
    }
    for (int ch = 0; ch < nch; ch++) orc_coset_ifft_inplace(&vals[(size_t)ch * L], log_l, GL_GENERATOR);
    return vals;
}

/* ---- the proof object (flat, shared with the engine) ----
 * [wires_cap | zs_pp_cap | quotient_cap] (3 x 2^h x 4) | openings: constants, sigmas, wires, zs, zs_next, partial products,
 * quotient polys (each (a, b)) | FRI proof blob */
struct OrcProof {
    vec64 wires_cap, zs_pp_cap, quotient_cap;
    vec2 constants, sigmas, wires, zs, zs_next, pps, quotient;
    OrcFriProof fri;
};

static inline std::vector<OrcFriBatchInfo> orc_fri_instance(const OrcCircuit &C, gl2 zeta) {
    std::vector<OrcFriBatchInfo> inst(2);
    inst[0].point = zeta;
    for (int p = 0; p < C.num_constants() + C.num_routed; p++) inst[0].polys.push_back({0, p});
    for (int p = 0; p < C.num_wires; p++) inst[0].polys.push_back({1, p});
    for (int p = 0; p < C.num_zs_pp(); p++) inst[0].polys.push_back({2, p});
    for (int p = 0; p < C.num_quotient_polys(); p++) inst[0].polys.push_back({3, p});
    inst[1].point = gl2_scale(zeta, gl_root_of_unity(C.degree_bits));
    for (int p = 0; p < C.num_challenges; p++) inst[1].polys.push_back({2, p});
    return inst;
}
static inline OrcFriParams orc_fri_params(const OrcCircuit &C) {
    OrcFriParams P;
    P.degree_bits = C.degree_bits; P.rate_bits = C.rate_bits; P.cap_height = C.cap_height; P.pow_bits = C.pow_bits;
    P.num_query_rounds = C.num_query_rounds;
    P.arity_bits = orc_fri_reduction_arity_bits(C.degree_bits, C.rate_bits, C.cap_height, 4, 5);
    return P;
}

/* prove_with_partition_witness after witness generation.  cs = the constants||sigmas batch from build();
 * wire_values [num_wires][n]; sigma_values [num_routed][n]; public_inputs_hash given.  Returns false when the
 * quotient is not a polynomial of degree < 8n ("Quotient has failed...": cannot happen by length here) -- kept for
 * signature symmetry. */
static inline bool orc_prove(const OrcCircuit &C, const OrcBatch &cs, const u64 *wire_values, const u64 *sigma_values,
                             const u64 pi_hash[4], OrcProof &proof) {
    const size_t n = (size_t)1 << C.degree_bits;
    const int nch = C.num_challenges;
    OrcChallenger ch; orc_ch_init(&ch);
    orc_ch_observe_n(&ch, C.circuit_digest, 4);
    orc_ch_observe_n(&ch, pi_hash, 4);
    OrcBatch wires; wires.num_polys = C.num_wires; wires.degree_log = C.degree_bits; wires.rate_bits = C.rate_bits;
    if (orc_batch_from_values(wires, wire_values, C.cap_height)) return false;
    orc_observe_cap(&ch, wires.tree.cap);
    u64 betas[8], gammas[8], alphas[8];
    for (int i = 0; i < nch; i++) betas[i] = orc_ch_challenge(&ch);
    for (int i = 0; i < nch; i++) gammas[i] = orc_ch_challenge(&ch);
    vec64 zpp_vals = orc_partial_products(C, wire_values, sigma_values, betas, gammas);
    OrcBatch zs_pp; zs_pp.num_polys = C.num_zs_pp(); zs_pp.degree_log = C.degree_bits; zs_pp.rate_bits = C.rate_bits;
    if (orc_batch_from_values(zs_pp, zpp_vals.data(), C.cap_height)) return false;
    orc_observe_cap(&ch, zs_pp.tree.cap);
    for (int i = 0; i < nch; i++) alphas[i] = orc_ch_challenge(&ch);
    vec64 q = orc_quotient_polys(C, cs, wires, zs_pp, pi_hash, betas, gammas, alphas);
    if (q.empty()) return false;
    /* trim_to_len(8n) is the identity (length check); chunks(n): chunk k of challenge c is polynomial c*8+k */
    OrcBatch quot; quot.num_polys = C.num_quotient_polys(); quot.degree_log = C.degree_bits; quot.rate_bits = C.rate_bits;
    quot.coeffs = q;   /* [nch][8n] contiguous == [nch*8][n] */
    if (orc_batch_from_coeffs(quot, C.cap_height)) return false;
    orc_observe_cap(&ch, quot.tree.cap);
    gl2 zeta = orc_ch_challenge_ext(&ch);
    gl2 gzeta = gl2_scale(zeta, gl_root_of_unity(C.degree_bits));
    vec2 e_cs = orc_batch_eval(cs, zeta), e_w = orc_batch_eval(wires, zeta), e_z = orc_batch_eval(zs_pp, zeta);
    vec2 e_zn = orc_batch_eval(zs_pp, gzeta), e_q = orc_batch_eval(quot, zeta);
    proof.wires_cap = wires.tree.cap; proof.zs_pp_cap = zs_pp.tree.cap; proof.quotient_cap = quot.tree.cap;
    proof.constants.assign(e_cs.begin(), e_cs.begin() + C.num_constants());
    proof.sigmas.assign(e_cs.begin() + C.num_constants(), e_cs.end());
    proof.wires = e_w;
    proof.zs.assign(e_z.begin(), e_z.begin() + nch);
    proof.zs_next.assign(e_zn.begin(), e_zn.begin() + nch);
    proof.pps.assign(e_z.begin() + nch, e_z.end());
    proof.quotient = e_q;
    /* observe_openings: [constants, sigmas, wires, zs, partial products, quotient] then [zs_next] */
    for (const vec2 *v : {&proof.constants, &proof.sigmas, &proof.wires, &proof.zs, &proof.pps, &proof.quotient, &proof.zs_next})
        for (const gl2 &e : *v) orc_ch_observe_ext(&ch, e);
    proof.fri = orc_fri_prove(orc_fri_instance(C, zeta), {&cs, &wires, &zs_pp, &quot}, &ch, orc_fri_params(C));
    (void)n;
    return true;
}

/* verify_with_challenges: recompute the challenges, check vanishing(zeta) == Z_H(zeta) * quotient(zeta) per challenge
 * (quotient recombined from its chunks with powers of zeta^n), then the FRI proof.  0 = accept. */
static inline int orc_verify(const OrcCircuit &C, const vec64 &cs_cap, const u64 pi_hash[4], const OrcProof &proof) {
    const int nch = C.num_challenges;
    OrcChallenger ch; orc_ch_init(&ch);
    orc_ch_observe_n(&ch, C.circuit_digest, 4);
    orc_ch_observe_n(&ch, pi_hash, 4);
    orc_observe_cap(&ch, proof.wires_cap);
    u64 betas[8], gammas[8], alphas[8];
    for (int i = 0; i < nch; i++) betas[i] = orc_ch_challenge(&ch);
    for (int i = 0; i < nch; i++) gammas[i] = orc_ch_challenge(&ch);
    orc_observe_cap(&ch, proof.zs_pp_cap);
    for (int i = 0; i < nch; i++) alphas[i] = orc_ch_challenge(&ch);
    orc_observe_cap(&ch, proof.quotient_cap);
    gl2 zeta = orc_ch_challenge_ext(&ch);
    if ((int)proof.constants.size() != C.num_constants() || (int)proof.sigmas.size() != C.num_routed || (int)proof.wires.size() != C.num_wires ||
        (int)proof.zs.size() != nch || (int)proof.zs_next.size() != nch || (int)proof.pps.size() != nch * C.num_partial_products() ||
        (int)proof.quotient.size() != C.num_quotient_polys()) return 20;
    for (const vec2 *v : {&proof.constants, &proof.sigmas, &proof.wires, &proof.zs, &proof.pps, &proof.quotient, &proof.zs_next})
        for (const gl2 &e : *v) orc_ch_observe_ext(&ch, e);
    /* vanishing polynomial at zeta over the extension field */
    const u64 n = (u64)1 << C.degree_bits;
    gl2 zeta_pow_n = gl2_pow(zeta, n);
    gl2 z_h = gl2_sub(zeta_pow_n, gl2_from(1));
    gl2 l0 = gl2_mul(z_h, gl2_inv(gl2_scale(gl2_sub(zeta, gl2_from(1)), n % GL_P)));   /* eval_l_0(n, zeta) */
    gl2 res[8];
    orc_eval_vanishing<OpsExt>(C, zeta, l0, proof.constants.data(), proof.sigmas.data(), proof.wires.data(), proof.zs.data(),
                               proof.zs_next.data(), proof.pps.data(), pi_hash, betas, gammas, alphas, res);
    for (int c = 0; c < nch; c++) {
        gl2 q = gl2_from(0);   /* reduce_with_powers(chunk evals, zeta^n) */
        for (int k = C.quotient_degree_factor; k-- > 0;) q = gl2_add(gl2_mul(q, zeta_pow_n), proof.quotient[c * C.quotient_degree_factor + k]);
        if (!gl2_eq(res[c], gl2_mul(z_h, q))) return 21;
    }
    std::vector<OrcFriBatchInfo> inst = orc_fri_instance(C, zeta);
    std::vector<vec2> ops(2);
    for (const vec2 *v : {&proof.constants, &proof.sigmas, &proof.wires, &proof.zs, &proof.pps, &proof.quotient}) ops[0].insert(ops[0].end(), v->begin(), v->end());
    ops[1] = proof.zs_next;
    std::vector<vec64> caps = {cs_cap, proof.wires_cap, proof.zs_pp_cap, proof.quotient_cap};
    std::vector<int> lens = {C.num_constants() + C.num_routed, C.num_wires, C.num_zs_pp(), C.num_quotient_polys()};
    return orc_fri_verify(inst, ops, caps, lens, proof.fri, &ch, orc_fri_params(C));
}

static inline vec64 orc_proof_blob(const OrcProof &p) {
    vec64 o;
    for (const vec64 *c : {&p.wires_cap, &p.zs_pp_cap, &p.quotient_cap}) o.insert(o.end(), c->begin(), c->end());
    for (const vec2 *v : {&p.constants, &p.sigmas, &p.wires, &p.zs, &p.zs_next, &p.pps, &p.quotient})
        for (const gl2 &e : *v) { o.push_back(e.a); o.push_back(e.b); }
    vec64 f = orc_fri_proof_blob(p.fri);
    o.insert(o.end(), f.begin(), f.end());
    return o;
}
static inline bool orc_proof_from_blob(const OrcCircuit &C, const u64 *b, size_t len, OrcProof &p) {
    size_t cap = (size_t)4 << C.cap_height, i = 0;
    size_t nop = C.num_constants() + C.num_routed + C.num_wires + 2 * C.num_challenges + C.num_challenges * C.num_partial_products() + C.num_quotient_polys();
    if (len < 3 * cap + 2 * nop) return false;
    p.wires_cap.assign(b, b + cap); p.zs_pp_cap.assign(b + cap, b + 2 * cap); p.quotient_cap.assign(b + 2 * cap, b + 3 * cap);
    i = 3 * cap;
    auto rd = [&](vec2 &v, int k) { for (int j = 0; j < k; j++, i += 2) v.push_back(gl2_make(b[i], b[i + 1])); };
    rd(p.constants, C.num_constants()); rd(p.sigmas, C.num_routed); rd(p.wires, C.num_wires); rd(p.zs, C.num_challenges);
    rd(p.zs_next, C.num_challenges); rd(p.pps, C.num_challenges * C.num_partial_products()); rd(p.quotient, C.num_quotient_polys());
    return orc_fri_proof_from_blob(b + i, len - i, p.fri);
}
#endif
