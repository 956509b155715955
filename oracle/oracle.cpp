/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).  C ABI used by tests/ (ctypes) and by bench.py's
 * cpu_baseline / --impl reference legs.  Never loaded by the product package. */
#include <chrono>
#include <cstdio>
#include <omp.h>
#include "commit.h"
#include "challenger.h"
#include "fri.h"
#include "plonk.h"

double orc_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

void orc_poseidon_permute(u64 *state, int naive) { if (naive) orc_poseidon_naive(state); else orc_poseidon(state); }
void orc_hash_n(const u64 *in, size_t n, u64 *out, int or_noop) { if (or_noop) orc_hash_or_noop(in, n, out); else orc_hash_no_pad(in, n, out); }
void orc_compress(const u64 *l, const u64 *r, u64 *out) { orc_two_to_one(l, r, out); }

u64 orc_gl_mul(u64 a, u64 b) { return gl_mul(gl_canon(a), gl_canon(b)); }
u64 orc_gl_inv(u64 a) { return gl_inv(gl_canon(a)); }
u64 orc_gl_pow(u64 a, u64 e) { return gl_pow(gl_canon(a), e); }
u64 orc_root_of_unity(int k) { return gl_root_of_unity(k); }

void orc_fft(u64 *a, int log_n) { size_t n = (size_t)1 << log_n; for (size_t i = 0; i < n; i++) a[i] = gl_canon(a[i]); orc_fft_inplace(a, log_n); }
void orc_ifft(u64 *a, int log_n) { size_t n = (size_t)1 << log_n; for (size_t i = 0; i < n; i++) a[i] = gl_canon(a[i]); orc_ifft_inplace(a, log_n); }
void orc_lde(const u64 *coeffs, int log_n, int rate_bits, u64 shift, u64 *out) { orc_coset_lde(coeffs, log_n, rate_bits, shift, out); }

/* ---- PolynomialBatch handle ---- */
void *orc_batch_from_values_c(const u64 *values, int C, int log_n, int rate_bits, int cap_height) {
    OrcBatch *b = new OrcBatch();
    b->num_polys = C; b->degree_log = log_n; b->rate_bits = rate_bits;
    if (orc_batch_from_values(*b, values, cap_height)) { delete b; return nullptr; }
    return b;
}
void *orc_batch_from_coeffs_c(const u64 *coeffs, int C, int log_n, int rate_bits, int cap_height) {
    OrcBatch *b = new OrcBatch();
    b->num_polys = C; b->degree_log = log_n; b->rate_bits = rate_bits;
    size_t n = (size_t)1 << log_n;
    b->coeffs.resize((size_t)C * n);
    for (size_t i = 0; i < (size_t)C * n; i++) b->coeffs[i] = gl_canon(coeffs[i]);
    if (orc_batch_from_coeffs(*b, cap_height)) { delete b; return nullptr; }
    return b;
}
void orc_batch_free(void *h) { delete (OrcBatch *)h; }
const u64 *orc_batch_coeffs(void *h) { return ((OrcBatch *)h)->coeffs.data(); }
const u64 *orc_batch_leaves(void *h) { return ((OrcBatch *)h)->tree.leaves.data(); }
const u64 *orc_batch_digests(void *h) { return ((OrcBatch *)h)->tree.digests.data(); }
size_t orc_batch_num_digests(void *h) { return ((OrcBatch *)h)->tree.digests.size() / 4; }
const u64 *orc_batch_cap(void *h) { return ((OrcBatch *)h)->tree.cap.data(); }
int orc_batch_prove(void *h, size_t leaf_index, u64 *siblings) { return orc_merkle_prove(((OrcBatch *)h)->tree, leaf_index, siblings); }
void orc_batch_times(void *h, double *out4) {
    OrcBatch *b = (OrcBatch *)h;
    out4[0] = b->t_ifft; out4[1] = b->t_lde; out4[2] = b->t_transpose; out4[3] = b->t_tree;
}

/* ---- bare MerkleTree::new over row-major leaves ---- */
void *orc_merkle_new(const u64 *leaves, size_t num_leaves, size_t leaf_len, int cap_height) {
    OrcBatch *b = new OrcBatch();
    b->tree.num_leaves = num_leaves; b->tree.leaf_len = leaf_len; b->tree.cap_height = cap_height;
    b->tree.leaves.resize(num_leaves * leaf_len);
    for (size_t i = 0; i < num_leaves * leaf_len; i++) b->tree.leaves[i] = gl_canon(leaves[i]);
    if (orc_merkle_build(b->tree)) { delete b; return nullptr; }
    return b;
}
int orc_merkle_verify_c(const u64 *leaf, size_t leaf_len, size_t leaf_index, const u64 *cap, const u64 *siblings, int n) {
    return orc_merkle_verify(leaf, leaf_len, leaf_index, cap, siblings, n);
}

/* ---- Challenger ---- */
void *orc_challenger_new() { OrcChallenger *c = new OrcChallenger(); orc_ch_init(c); return c; }
void orc_challenger_free(void *c) { delete (OrcChallenger *)c; }
void orc_challenger_observe(void *c, const u64 *e, size_t n) { orc_ch_observe_n((OrcChallenger *)c, e, n); }
u64 orc_challenger_get(void *c) { return orc_ch_challenge((OrcChallenger *)c); }


/* ---- openings + FRI ----
 * instance blob: [num_batches] then per batch [point.a, point.b, num_polys, (oracle << 32 | poly) x num_polys]
 * params: [degree_bits, rate_bits, cap_height, pow_bits, num_query_rounds, num_arity, arity_bits...] */
static std::vector<OrcFriBatchInfo> parse_instance(const u64 *b) {
    std::vector<OrcFriBatchInfo> out;
    size_t i = 0;
    u64 nb = b[i++];
    for (u64 k = 0; k < nb; k++) {
        OrcFriBatchInfo bi;
        bi.point = gl2_make(b[i], b[i + 1]); i += 2;
        u64 np = b[i++];
        for (u64 j = 0; j < np; j++) { OrcFriPolyRef r = {(int)(b[i] >> 32), (int)(b[i] & 0xffffffffu)}; bi.polys.push_back(r); i++; }
        out.push_back(bi);
    }
    return out;
}
static OrcFriParams parse_params(const int *p) {
    OrcFriParams P;
    P.degree_bits = p[0]; P.rate_bits = p[1]; P.cap_height = p[2]; P.pow_bits = p[3]; P.num_query_rounds = p[4];
    for (int i = 0; i < p[5]; i++) P.arity_bits.push_back(p[6 + i]);
    return P;
}
// Horner evaluation of one coefficient vector at a base-field point (PolynomialCoeffs::eval); used by the size-independent
// property tests at sizes where a whole oracle commit would take minutes
u64 orc_poly_eval(const u64 *coeffs, u64 n, u64 x) {
    u64 acc = 0;
    x = gl_canon(x);
    for (u64 i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), gl_canon(coeffs[i]));
    return gl_canon(acc);
}
void orc_batch_eval_c(void *h, const u64 *z, u64 *out) {
    vec2 v = orc_batch_eval(*(OrcBatch *)h, gl2_make(gl_canon(z[0]), gl_canon(z[1])));
    for (size_t i = 0; i < v.size(); i++) { out[2 * i] = v[i].a; out[2 * i + 1] = v[i].b; }
}
int orc_fri_arity_bits_c(int degree_bits, int rate_bits, int cap_height, int arity_bits, int final_poly_bits, int *out) {
    std::vector<int> r = orc_fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits, final_poly_bits);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
    return (int)r.size();
}
void *orc_fri_prove_c(const u64 *instance, void **oracles, int num_oracles, void *challenger, const int *params) {
    std::vector<const OrcBatch *> os;
    for (int i = 0; i < num_oracles; i++) os.push_back((const OrcBatch *)oracles[i]);
    OrcFriProof p = orc_fri_prove(parse_instance(instance), os, (OrcChallenger *)challenger, parse_params(params));
    return new vec64(orc_fri_proof_blob(p));
}
size_t orc_blob_len(void *b) { return ((vec64 *)b)->size(); }
const u64 *orc_blob_data(void *b) { return ((vec64 *)b)->data(); }
void orc_blob_free(void *b) { delete (vec64 *)b; }
/* openings: per batch, the claimed values in instance order, flattened [(a, b)...]; caps: num_oracles x (2^h * 4) */
int orc_fri_verify_c(const u64 *instance, const u64 *openings, const u64 *caps, const int *leaf_lens, int num_oracles,
                     const u64 *proof_blob, size_t proof_len, void *challenger, const int *params) {
    std::vector<OrcFriBatchInfo> inst = parse_instance(instance);
    OrcFriParams P = parse_params(params);
    std::vector<vec2> ops;
    size_t i = 0;
    for (const OrcFriBatchInfo &b : inst) {
        vec2 v;
        for (size_t j = 0; j < b.polys.size(); j++) { v.push_back(gl2_make(openings[i], openings[i + 1])); i += 2; }
        ops.push_back(v);
    }
    std::vector<vec64> cps;
    std::vector<int> lens;
    size_t capsz = (size_t)4 << P.cap_height;
    for (int o = 0; o < num_oracles; o++) { cps.emplace_back(caps + o * capsz, caps + (o + 1) * capsz); lens.push_back(leaf_lens[o]); }
    OrcFriProof p;
    if (!orc_fri_proof_from_blob(proof_blob, proof_len, p)) return 100;
    return orc_fri_verify(inst, ops, cps, lens, p, (OrcChallenger *)challenger, P);
}
void orc_challenger_state(void *c, u64 *out /* 12 state + 1 n_in + 8 in + 1 n_out + 8 out */) {
    OrcChallenger *ch = (OrcChallenger *)c;
    memcpy(out, ch->state, 96); out[12] = ch->n_in; memcpy(out + 13, ch->in_buf, 64); out[21] = ch->n_out; memcpy(out + 22, ch->out_buf, 64);
}


/* ---- plonk rows (a5, a6) and the whole prove / verify ----
 * circuit blob: [degree_bits, num_wires, num_routed, num_gate_constants, num_selectors, num_challenges,
 *   quotient_degree_factor, rate_bits, cap_height, pow_bits, num_query_rounds, num_gates,
 *   (kind, selector_index, group_start, group_end) x num_gates, circuit_digest x 4] */
/* Both layouts of include/plonky2_b200.h.  From a version-2 description the oracle reads the gate KINDS and parameters only:
 * the bytecode programs in it are what the engine runs, the oracle evaluates the gates from its own formulas (gates.h). */
void *orc_circuit_new(const u64 *b0) {
    OrcCircuit *c = new OrcCircuit();
    const bool v2 = b0[0] == 0x32424B4C50ULL;
    const u64 *b = v2 ? b0 + 2 : b0;
    c->degree_bits = (int)b[0]; c->num_wires = (int)b[1]; c->num_routed = (int)b[2]; c->num_gate_constants = (int)b[3];
    c->num_selectors = (int)b[4]; c->num_challenges = (int)b[5]; c->quotient_degree_factor = (int)b[6]; c->rate_bits = (int)b[7];
    c->cap_height = (int)b[8]; c->pow_bits = (int)b[9]; c->num_query_rounds = (int)b[10];
    int ng = (int)b[11];
    for (int i = 0; i < ng; i++) {
        const u64 *e = v2 ? b0 + 20 + 12 * i : b + 12 + 4 * i;
        OrcGateInfo g = {(int)e[0], (int)e[1], (int)e[2], (int)e[3], {0, 0, 0, 0}};
        if (v2) for (int k = 0; k < 4; k++) g.p[k] = (int)e[4 + k];
        else { if (g.kind == ORC_G_CONSTANT) g.p[0] = 2; if (g.kind == ORC_G_ARITHMETIC) g.p[0] = 20; }
        c->gates.push_back(g);
    }
    for (int i = 0; i < 4; i++) c->circuit_digest[i] = gl_canon(v2 ? b0[16 + i] : b[12 + 4 * ng + i]);
    u64 k = 1;
    for (int j = 0; j < c->num_routed; j++) { c->k_is.push_back(k); k = gl_mul(k, GL_GENERATOR); }
    return c;
}
void orc_circuit_free(void *c) { delete (OrcCircuit *)c; }
void orc_partial_products_c(void *c, const u64 *wires, const u64 *sigmas, const u64 *betas, const u64 *gammas, u64 *out) {
    vec64 r = orc_partial_products(*(OrcCircuit *)c, wires, sigmas, betas, gammas);
    memcpy(out, r.data(), r.size() * 8);
}
int orc_quotient_c(void *c, void *cs, void *wires, void *zs_pp, const u64 *pi_hash, const u64 *betas, const u64 *gammas, const u64 *alphas, u64 *out) {
    vec64 r = orc_quotient_polys(*(OrcCircuit *)c, *(OrcBatch *)cs, *(OrcBatch *)wires, *(OrcBatch *)zs_pp, pi_hash, betas, gammas, alphas);
    if (r.empty()) return 1;
    memcpy(out, r.data(), r.size() * 8);
    return 0;
}
void *orc_prove_c(void *c, void *cs, const u64 *wire_values, const u64 *sigma_values, const u64 *pi_hash) {
    OrcProof p;
    if (!orc_prove(*(OrcCircuit *)c, *(OrcBatch *)cs, wire_values, sigma_values, pi_hash, p)) return nullptr;
    return new vec64(orc_proof_blob(p));
}
int orc_verify_c(void *c, const u64 *cs_cap, const u64 *pi_hash, const u64 *blob, size_t len) {
    OrcCircuit *C = (OrcCircuit *)c;
    OrcProof p;
    if (!orc_proof_from_blob(*C, blob, len, p)) return 100;
    vec64 cap(cs_cap, cs_cap + ((size_t)4 << C->cap_height));
    return orc_verify(*C, cap, pi_hash, p);
}

} /* extern "C" */
