// CPU replay of the CUDA kernels -- TEST HARNESS ONLY (never shipped, never loaded by the product package).
//
// The development container has no GPU, so this file compiles the kernel BODIES of
// eth-lc-plonky2_b200/csrc/{ntt,merkle,poseidon}.cuh (written as __host__ __device__ functions) with g++ and
// replays every launch of a plan block by block, thread by thread, phase by phase (each __syncthreads()
// boundary is a loop boundary).  It checks the index math, table construction and planner on the CPU against
// the oracle before GPU minutes are spent; the `-m gpu` tests then check the real kernels through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
// ranges seen by poseidon_f64.cuh's tracking hooks: [0] min and [1] max entering a fold, [2] largest |x| entering a renorm
static double g_pf_range[3] = {0.0, 0.0, 0.0};
#define PF_TRACK_FOLD(x) do { double v__ = (x); if (v__ < g_pf_range[0]) g_pf_range[0] = v__; if (v__ > g_pf_range[1]) g_pf_range[1] = v__; } while (0)
#define PF_TRACK_RENORM(x) do { double a__ = std::fabs(x); if (a__ > g_pf_range[2]) g_pf_range[2] = a__; } while (0)
static double g_l3_maxabs = 0.0;   // largest lazy 96-bit magnitude built by ntt.cuh's l3_* helpers (L3_TRACK hook)
#define L3_TRACK(v) do { double a__ = std::fabs((double)(v)); if (a__ > g_l3_maxabs) g_l3_maxabs = a__; } while (0)
#include "../../eth-lc-plonky2_b200/csrc/merkle.cuh"
#include "../../eth-lc-plonky2_b200/csrc/ntt.cuh"
#include "../../eth-lc-plonky2_b200/csrc/ntt_plan.h"
#include "../../eth-lc-plonky2_b200/csrc/plonk.cuh"

static std::vector<std::unique_ptr<std::vector<u64>>> g_tables;

static NttTableStore make_store() {
    NttTableStore ts;
    ts.upload = [](const std::vector<u64> &t) -> const u64 * {
        g_tables.emplace_back(new std::vector<u64>(t));
        return g_tables.back()->data();
    };
    return ts;
}

// every round is a __syncthreads() interval: run it for all threads before the next one starts
template <int R, bool LAST>
static void replay_round(const NttPass &p, u64 *sm, u32 t0, u32 threads) {
    for (u32 t = 0; t < threads; t++) ntt_round<R, LAST>(p, sm, p.tw_local, t0, t, threads);
}
template <bool LAST, bool INV>
static void replay_round16(const NttPass &p, u64 *sm, const u64 *tau, u32 t0, u32 threads) {
    for (u32 t = 0; t < threads; t++) ntt_round16<LAST, INV>(p, sm, tau, t0, t, threads);
}
template <bool INV>
static void replay_rounds(const NttPass &p, u64 *sm, u32 threads) {
    u32 t0 = 0;
    const u32 rem = p.log_p & 3;
    if (rem == 1) { if (p.log_p == 1) replay_round<1, true>(p, sm, t0, threads); else replay_round<1, false>(p, sm, t0, threads); t0 += 1; }
    if (rem == 2) { if (p.log_p == 2) replay_round<2, true>(p, sm, t0, threads); else replay_round<2, false>(p, sm, t0, threads); t0 += 2; }
    if (rem == 3) { if (p.log_p == 3) replay_round<3, true>(p, sm, t0, threads); else replay_round<3, false>(p, sm, t0, threads); t0 += 3; }
    const u64 *tau = p.tau_tab;          // same walk through the per-round tables as NTT_ROUNDS
    for (; t0 + 4 < p.log_p; t0 += 4) { replay_round16<false, INV>(p, sm, tau, t0, threads); tau += 15u << (p.log_p - t0 - 4); }
    if (t0 < p.log_p) replay_round16<true, INV>(p, sm, tau, t0, threads);
}

template <int MODE>
static void replay_mode(const NttLaunch &l, u64 grid) {
    std::vector<u64> sm(l.smem / 8 + 16);
    const NttPass &p = l.p;
    for (u64 bid = 0; bid < grid; bid++) {
        for (u64 tile_i = bid; tile_i < p.num_tiles; tile_i += grid) {
            u64 tile = tile_i + p.tile_rot;       // same walk as ntt_pass_kernel
            if (tile >= p.num_tiles) tile -= p.num_tiles;
            for (u32 t = 0; t < l.threads; t++) ntt_load<MODE>(p, sm.data(), tile, t, l.threads);
            replay_rounds<(MODE >= NTT_INTT_P1)>(p, sm.data(), l.threads);
            for (u32 t = 0; t < l.threads; t++) ntt_store<MODE>(p, sm.data(), tile, t, l.threads);
        }
    }
}
static void replay(const std::vector<NttLaunch> &plan, u64 grid) {
    for (const NttLaunch &l : plan) {
        switch (l.mode) {
            case NTT_LDE_FIRST: replay_mode<NTT_LDE_FIRST>(l, grid); break;
            case NTT_LDE_SINGLE: replay_mode<NTT_LDE_SINGLE>(l, grid); break;
            case NTT_DIF_LAST: replay_mode<NTT_DIF_LAST>(l, grid); break;
            case NTT_INTT_P1: replay_mode<NTT_INTT_P1>(l, grid); break;
            case NTT_INTT_P2: replay_mode<NTT_INTT_P2>(l, grid); break;
            case NTT_INTT_SINGLE: replay_mode<NTT_INTT_SINGLE>(l, grid); break;
            default: abort();
        }
    }
}

extern "C" {
int emu_batch_sharded(const u64 *values, u32 C, u32 log_n, u32 rate_bits, u32 cap_height, int is_values, u64 grid,
                      u32 log_shards, u64 *coeffs, u64 *lde, u64 *digests, u64 *cap);

void emu_poseidon_permute(const u64 *in, u64 *out, size_t count) {
    for (size_t i = 0; i < count; i++) {
        u64 s[12];
        for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
        poseidon_permute(s);
        for (int k = 0; k < 12; k++) out[12 * i + k] = gl_canon(s[k]);
    }
}

u64 emu_gl_mul(u64 a, u64 b) { return gl_canon(gl_mul(a, b)); }
u64 emu_gl_add(u64 a, u64 b) { return gl_canon(gl_add(a, b)); }
u64 emu_gl_sub(u64 a, u64 b) { return gl_canon(gl_sub(a, b)); }

// values [C][n] -> coeffs [C][n], lde [C][L] (column-major, bit-reversed rows), digests, cap.  Returns 0 on success.
int emu_batch_from_values(const u64 *values, u32 C, u32 log_n, u32 rate_bits, u32 cap_height, int is_values, u64 grid,
                          u64 *coeffs, u64 *lde, u64 *digests, u64 *cap) {
    return emu_batch_sharded(values, C, log_n, rate_bits, cap_height, is_values, grid, 0, coeffs, lde, digests, cap);
}

// same with the LDE written as 2^log_shards row shards [G][C][L/G]; the tree is only built for log_shards = 0
int emu_batch_sharded(const u64 *values, u32 C, u32 log_n, u32 rate_bits, u32 cap_height, int is_values, u64 grid,
                      u32 log_shards, u64 *coeffs, u64 *lde, u64 *digests, u64 *cap) {
    NttTableStore ts = make_store();
    const u64 n = (u64)1 << log_n, L = n << rate_bits;
    std::vector<NttLaunch> plan;
    if (is_values) {
        if (!ntt_plan_intt(ts, values, n, lde, n, coeffs, n, C, log_n, plan)) return 1;
        replay(plan, grid);
    } else {
        memcpy(coeffs, values, (size_t)C * n * 8);
    }
    plan.clear();
    if (!ntt_plan_lde(ts, coeffs, n, lde, C, log_n, rate_bits, log_shards, plan)) return 2;
    replay(plan, grid);
    if (log_shards) { g_tables.clear(); return 0; }
    MerkleParams mp;
    mp.data = lde; mp.row_stride = 1; mp.col_stride = L; mp.width = C; mp.noop_max = 4; mp.num_leaves = L;
    mp.num_layers = log_n + rate_bits - cap_height; mp.digests = digests; mp.cap = cap;
    for (u64 j = 0; j < L; j++) merkle_hash_leaf(mp, j);
    for (u32 layer = 0; layer < mp.num_layers; layer++)
        for (u64 gidx = 0; gidx < (L >> (layer + 1)); gidx++) merkle_hash_node(mp, layer, gidx);
    g_tables.clear();
    return 0;
}

// Fused exchange (eng_lde_peer_dev): `rank` transforms its C_r columns and the last pass stores row shard g through
// shard_out[g] = mats[g] + col0 * L/G, mats[g] being "rank g's" leaf matrix [C_total][L/G]; tiles are walked from shard
// `rank` on.  Returns 0 on success.
int emu_lde_peer(const u64 *values, u32 C_r, u32 log_n, u32 rate_bits, u64 grid, u32 log_shards, u32 rank, u32 col0,
                 u64 *const *mats, u64 *coeffs, u64 *scratch) {
    NttTableStore ts = make_store();
    const u64 n = (u64)1 << log_n, L = n << rate_bits;
    std::vector<NttLaunch> plan;
    if (!ntt_plan_intt(ts, values, n, scratch, n, coeffs, n, C_r, log_n, plan)) return 1;
    replay(plan, grid);
    plan.clear();
    u64 *so[NTT_MAX_SHARDS];
    for (u32 g = 0; g < (1u << log_shards); g++) so[g] = mats[g] + (size_t)col0 * (L >> log_shards);
    if (!ntt_plan_lde(ts, coeffs, n, scratch, C_r, log_n, rate_bits, log_shards, plan, so, rank)) return 2;
    replay(plan, grid);
    g_tables.clear();
    return 0;
}

// describes the launch list (for tests of the planner): fills up to max entries of [mode, log_p, log_a, threads, smem, tiles]
int emu_plan(u32 C, u32 log_n, u32 rate_bits, int intt, u64 *out, int max) {
    NttTableStore ts = make_store();
    std::vector<NttLaunch> plan;
    std::vector<u64> dummy(1);
    bool ok = intt ? ntt_plan_intt(ts, dummy.data(), 0, dummy.data(), 0, dummy.data(), 0, C, log_n, plan)
                   : ntt_plan_lde(ts, dummy.data(), 0, dummy.data(), C, log_n, rate_bits, 0, plan);
    g_tables.clear();
    if (!ok) return -1;
    int k = 0;
    for (const NttLaunch &l : plan) {
        if (k >= max) break;
        u64 *o = out + 6 * k++;
        o[0] = l.mode; o[1] = l.p.log_p; o[2] = l.p.log_a; o[3] = l.threads; o[4] = l.smem; o[5] = l.p.num_tiles;
    }
    return k;
}

// ---- plonk rows: replay of the quotient kernels' bodies (quot_perm_point / quot_poseidon_point / quot_gates_point /
// quot_finish_point, l0_table_group) and of pp_row / pp_finish on host arrays ----
struct EmuCircuit {
    u32 degree_bits, num_wires, num_routed, num_gate_constants, num_selectors, num_challenges, qdf, qdb, rate_bits;
    std::vector<PlkGateDev> gates;
    std::vector<u64> prog, imm;
    u32 max_constraints = 0;
    int poseidon_index = -1;
};
// both layouts of include/plonky2_b200.h; the programs are rebuilt from kinds + parameters with gate_lib.h unless the
// description carries them
static bool emu_parse_circuit(const u64 *b, EmuCircuit &C) {
    const bool v2 = b[0] == 0x32424B4C50ull;
    const u64 *h = v2 ? b + 2 : b;
    C.degree_bits = (u32)h[0]; C.num_wires = (u32)h[1]; C.num_routed = (u32)h[2]; C.num_gate_constants = (u32)h[3];
    C.num_selectors = (u32)h[4]; C.num_challenges = (u32)h[5]; C.qdf = (u32)h[6]; C.rate_bits = (u32)h[7];
    C.qdb = 0;
    while ((1u << C.qdb) < C.qdf) C.qdb++;
    const u32 ng = (u32)h[11];
    GvmImmPool pool;
    const u64 *gt = v2 ? b + 20 : b + 12, *progs = gt + (size_t)ng * 12;
    if (v2) {
        const u64 *imms = progs + b[15];
        for (u64 i = GVM_NUM_PI; i < b[14]; i++) { pool.values.push_back(gl_canon(imms[i])); pool.index.emplace(gl_canon(imms[i]), (u32)i); }
    }
    for (u32 i = 0; i < ng; i++) {
        const u64 *e = gt + (size_t)i * (v2 ? 12 : 4);
        u32 p[4] = {0, 0, 0, 0};
        const u32 kind = (u32)e[0];
        if (v2) for (int k = 0; k < 4; k++) p[k] = (u32)e[4 + k];
        else { if (kind == PLK_CONSTANT) p[0] = 2; if (kind == PLK_ARITHMETIC) p[0] = 20; }
        std::vector<u64> program;
        if (v2 && e[10]) program.assign(progs + e[9], progs + e[9] + e[10]);
        else if (kind != PLK_NOOP) {
            GvmBuilder B(pool);
            if (!plk_build_gate(B, kind, p, C.num_wires, C.num_routed, C.num_gate_constants)) return false;
            program = B.finish();
            if (!B.ok) return false;
        }
        PlkGateDev g;
        memset(&g, 0, sizeof(g));
        g.prog_off = (u32)C.prog.size(); g.prog_len = (u32)program.size();
        g.selector_index = (u32)e[1]; g.group_start = (u32)e[2]; g.group_end = (u32)e[3]; g.row = i;
        g.kind = kind;
        for (int k = 0; k < 4; k++) g.p[k] = p[k];
        C.prog.insert(C.prog.end(), program.begin(), program.end());
        C.gates.push_back(g);
        if (kind == PLK_POSEIDON && C.poseidon_index < 0) C.poseidon_index = (int)i;
    }
    C.imm = pool.values;
    for (PlkGateDev &g : C.gates) {
        u32 nc = 0;
        if (g.prog_len && !gvm_validate(C.prog.data() + g.prog_off, g.prog_len, C.num_wires, C.num_gate_constants, (u32)C.imm.size(), &nc, nullptr)) return false;
        g.num_constraints = nc;
        if (nc > C.max_constraints) C.max_constraints = nc;
    }
    return true;
}
// LDEs are [cols][L] column-major in bit-reversed row order (the engine's layout); out: [num_challenges][Lq] natural order.
// native != 0: PoseidonGate through the FP64 evaluator (the default of the engine), otherwise through its bytecode.
int emu_quotient_values_sharded(const u64 *blob, const u64 *cs_lde, const u64 *wires_lde, const u64 *zs_lde, const u64 *pi_hash,
                                const u64 *betas, const u64 *gammas, const u64 *alphas, u64 *out, int native, u32 log_shards);
int emu_quotient_values(const u64 *blob, const u64 *cs_lde, const u64 *wires_lde, const u64 *zs_lde, const u64 *pi_hash,
                        const u64 *betas, const u64 *gammas, const u64 *alphas, u64 *out, int native) {
    return emu_quotient_values_sharded(blob, cs_lde, wires_lde, zs_lde, pi_hash, betas, gammas, alphas, out, native, 0);
}
// log_shards > 0: the multi-GPU prover's path -- every row shard is cut out of the LDEs as its own [cols][L/G] leaf matrix,
// evaluated with pos0 / count / by_position, and the gathered shards go through quot_unshard_point
int emu_quotient_values_sharded(const u64 *blob, const u64 *cs_lde, const u64 *wires_lde, const u64 *zs_lde, const u64 *pi_hash,
                                const u64 *betas, const u64 *gammas, const u64 *alphas, u64 *out, int native, u32 log_shards) {
    NttTableStore ts = make_store();
    static bool pf_built = false;   // the FP64 PoseidonGate evaluator reads the host tables
    if (!pf_built) { psd_f64_build_tables(h_pf); pf_built = true; }
    EmuCircuit C;
    if (!emu_parse_circuit(blob, C)) return 1;
    const u32 log_lq = C.degree_bits + C.qdb, nch = C.num_challenges;
    const u64 n = (u64)1 << C.degree_bits, Lq = (u64)1 << log_lq, L = n << C.rate_bits;
    const u32 npp = (C.num_routed + C.qdf - 1) / C.qdf - 1, perm_terms = nch + nch * (npp + 1);
    QuotParams q;
    memset(&q, 0, sizeof(q));
    q.log_n = C.degree_bits; q.log_lq = log_lq;
    q.num_wires = C.num_wires; q.num_routed = C.num_routed; q.num_selectors = C.num_selectors; q.num_gate_constants = C.num_gate_constants;
    q.num_challenges = nch; q.degree = C.qdf; q.npp = npp; q.stride = L; q.pos0 = 0; q.count = Lq;
    q.cs = cs_lde; q.wires = wires_lde; q.zs = zs_lde; q.out = out;
    q.k_is[0] = 1;
    for (int j = 1; j < PLK_MAX_ROUTED; j++) q.k_is[j] = h_gl_mul(q.k_is[j - 1], 7);
    u64 alpha_c[PLK_MAX_CHALLENGES] = {0, 0};
    for (u32 i = 0; i < nch; i++) { q.beta[i] = betas[i]; q.gamma[i] = gammas[i]; alpha_c[i] = alphas[i]; }
    q.apow_stride = perm_terms + C.max_constraints + 1; q.first_gate_term = perm_terms;
    std::vector<u64> apow_tab((size_t)PLK_MAX_CHALLENGES * q.apow_stride), acc((size_t)nch * Lq), l0(Lq);
    plk_fill_apow(alpha_c, nch, q.apow_stride, apow_tab.data());
    q.apow = apow_tab.data(); q.acc = acc.data(); q.l0 = l0.data();
    for (int i = 0; i < 4; i++) C.imm[i] = gl_canon(pi_hash[i]);
    // native bit 0: PoseidonGate through the FP64 evaluator; bit 1: the library gates through the compiled evaluators
    const bool nat = (native & 1) && C.poseidon_index >= 0, compiled = (native & 2) != 0;
    for (size_t i = 0; i < C.gates.size(); i++)
        C.gates[i].native = (int)i == C.poseidon_index ? 1 : (C.gates[i].kind < 32 && ((PLK_NATIVE_KINDS >> C.gates[i].kind) & 1)) ? 2 : 0;
    q.use_native_gates = compiled;
    q.gates = C.gates.data(); q.num_gates = (u32)C.gates.size(); q.prog = C.prog.data(); q.imm = C.imm.data();
    q.has_poseidon = nat;
    if (C.poseidon_index >= 0) q.poseidon = C.gates[C.poseidon_index];
    L0Params lp;
    memset(&lp, 0, sizeof(lp));
    lp.log_n = C.degree_bits; lp.log_lq = log_lq; lp.n_field = n % GL_P; lp.out = l0.data();
    const u64 g_pow_n = h_gl_pow(7, n), wq = h_gl_root_of_unity((int)C.qdb);
    for (u32 i = 0; i < (1u << C.qdb); i++) { lp.zh[i] = gl_canon(gl_sub(h_gl_mul(g_pow_n, h_gl_pow(wq, i)), 1)); q.zh_inv[i] = h_gl_inv(lp.zh[i]); }
    auto w = ts.w2((int)log_lq, false);
    q.w_lo = lp.w_lo = w.lo; q.w_hi = lp.w_hi = w.hi; q.w_lo_bits = lp.w_lo_bits = w.lo_bits;
    for (u64 grp = 0; grp * L0_BATCH < Lq; grp++) l0_table_group(lp, grp);
    if (log_shards == 0) {
        for (u64 pos = 0; pos < Lq; pos++) quot_perm_point(q, pos);
        if (nat) for (u64 pos = 0; pos < Lq; pos++) quot_poseidon_point(q, pos);
        if (compiled) for (u64 pos = 0; pos < Lq; pos++) quot_native_point<PLK_NATIVE_KINDS>(q, pos);
        for (u64 pos = 0; pos < Lq; pos++) quot_gates_point(q, pos);
        for (u64 pos = 0; pos < Lq; pos++) quot_finish_point(q, pos);
        g_tables.clear();
        return 0;
    }
    if (Lq != L || log_shards > C.qdb) { g_tables.clear(); return 2; }
    const u32 G = 1u << log_shards;
    const u64 count = Lq >> log_shards;
    const u32 ncs = C.num_selectors + C.num_gate_constants + C.num_routed, nzs = nch * (1 + npp);
    std::vector<u64> gathered((size_t)nch * Lq);
    for (u32 gsh = 0; gsh < G; gsh++) {
        auto cut = [&](const u64 *lde, u32 cols) {
            std::vector<u64> m((size_t)cols * count);
            for (u32 c2 = 0; c2 < cols; c2++) memcpy(&m[(size_t)c2 * count], lde + (size_t)c2 * L + (size_t)gsh * count, count * 8);
            return m;
        };
        std::vector<u64> mcs = cut(cs_lde, ncs), mw = cut(wires_lde, C.num_wires), mz = cut(zs_lde, nzs), sacc((size_t)nch * count);
        QuotParams s2 = q;
        s2.cs = mcs.data(); s2.wires = mw.data(); s2.zs = mz.data(); s2.stride = count; s2.pos0 = (u64)gsh * count; s2.count = count;
        s2.by_position = 1; s2.acc = sacc.data(); s2.out = &gathered[(size_t)gsh * nch * count];
        for (u64 t = 0; t < count; t++) quot_perm_point(s2, t);
        if (nat) for (u64 t = 0; t < count; t++) quot_poseidon_point(s2, t);
        if (compiled) for (u64 t = 0; t < count; t++) quot_native_point<PLK_NATIVE_KINDS>(s2, t);
        for (u64 t = 0; t < count; t++) quot_gates_point(s2, t);
        for (u64 t = 0; t < count; t++) quot_finish_point(s2, t);
    }
    for (u64 pos = 0; pos < Lq; pos++) quot_unshard_point(gathered.data(), out, log_lq, log_shards, nch, pos);
    g_tables.clear();
    return 0;
}
void emu_partial_products(const u64 *blob, const u64 *wires, const u64 *sigmas, const u64 *betas, const u64 *gammas, u64 *out) {
    NttTableStore ts = make_store();
    EmuCircuit C;
    if (!emu_parse_circuit(blob, C)) abort();
    PpParams p;
    memset(&p, 0, sizeof(p));
    p.log_n = C.degree_bits; p.num_routed = C.num_routed; p.num_challenges = C.num_challenges; p.degree = C.qdf;
    const u64 n = (u64)1 << C.degree_bits;
    std::vector<u64> row_prod((size_t)C.num_challenges * n);
    p.wires = wires; p.sigmas = sigmas; p.out = out; p.row_prod = row_prod.data();
    p.k_is[0] = 1;
    for (int j = 1; j < PLK_MAX_ROUTED; j++) p.k_is[j] = h_gl_mul(p.k_is[j - 1], 7);
    for (u32 i = 0; i < C.num_challenges; i++) { p.beta[i] = betas[i]; p.gamma[i] = gammas[i]; }
    auto w = ts.w2((int)C.degree_bits, false);
    p.w_lo = w.lo; p.w_hi = w.hi; p.w_lo_bits = w.lo_bits;
    for (u64 i = 0; i < n; i++) pp_row(p, i);
    for (u32 c = 0; c < C.num_challenges; c++) {   // the exclusive scan the three scan kernels perform
        u64 acc = 1;
        for (u64 i = 0; i < n; i++) { out[(u64)c * n + i] = gl_canon(acc); acc = gl_mul(acc, row_prod[(u64)c * n + i]); }
    }
    for (u64 i = 0; i < n; i++) pp_finish(p, i);
    g_tables.clear();
}
}

// the FP64 formulation of the permutation (poseidon_f64.cuh) replayed with host IEEE doubles
extern "C" void emu_poseidon_permute_f64(const u64 *in, u64 *out, size_t count) {
    static bool built = false;
    if (!built) { psd_f64_build_tables(h_pf); built = true; }
    for (size_t i = 0; i < count; i++) {
        u64 s[12];
        for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
        poseidon_permute_f64(s);
        for (int k = 0; k < 12; k++) out[12 * i + k] = gl_canon(s[k]);
    }
}

// ranges seen at the folds / re-normalisations since the last reset
extern "C" void emu_poseidon_f64_ranges(double *out, int reset) {
    for (int i = 0; i < 3; i++) { out[i] = g_pf_range[i]; if (reset) g_pf_range[i] = 0.0; }
}

extern "C" double emu_l3_maxabs(int reset) {
    double v = g_l3_maxabs;
    if (reset) g_l3_maxabs = 0.0;
    return v;
}
