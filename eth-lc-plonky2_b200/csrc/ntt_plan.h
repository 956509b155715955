// Host-side planning of NTT passes: tile shapes, twiddle / shift tables, launch lists.
// Pure host C++ (no CUDA calls): the engine uploads tables through NttTableStore::upload, the CPU replay
// harness in tests/emu keeps them in host memory.  See ntt.cuh for what each pass does.
#pragma once
#include <functional>
#include <map>
#include <tuple>
#include <vector>
#include "ntt.cuh"

struct NttLaunch {
    int mode;
    NttPass p;
    uint32_t threads;
    size_t smem;
};

struct NttTableStore {
    // copies a host table to wherever kernels read it from and returns that pointer (kept alive by the owner)
    std::function<const u64 *(const std::vector<u64> &)> upload;
    std::map<std::pair<int, int>, const u64 *> tw_local_cache;
    struct W2 { const u64 *lo, *hi; u32 lo_bits; };
    std::map<std::pair<int, int>, W2> w2_cache;
    struct Shift { const u64 *a, *b; };
    std::map<std::tuple<int, int, int>, Shift> shift_cache;

    // w_P^e (inverse: w_P^-e), e < P
    const u64 *tw_local(int log_p, bool inverse) {
        auto key = std::make_pair(log_p, (int)inverse);
        auto it = tw_local_cache.find(key);
        if (it != tw_local_cache.end()) return it->second;
        size_t half = (size_t)1 << log_p;
        std::vector<u64> t(half);
        u64 w = h_gl_root_of_unity(log_p);
        if (inverse) w = h_gl_inv(w);
        t[0] = 1;
        for (size_t i = 1; i < half; i++) t[i] = h_gl_mul(t[i - 1], w);
        const u64 *d = upload(t);
        if (!d) return nullptr;            // a failed upload is not cached (the owner reports it)
        return tw_local_cache[key] = d;
    }
    // tau tables of the non-final radix-16 rounds of a P-point network (layout: NttPass::tau_tab)
    std::map<std::pair<int, int>, const u64 *> tau_cache;
    const u64 *tau_tab(int log_p, bool inverse) {
        auto key = std::make_pair(log_p, (int)inverse);
        auto it = tau_cache.find(key);
        if (it != tau_cache.end()) return it->second;
        u64 w = h_gl_root_of_unity(log_p);
        if (inverse) w = h_gl_inv(w);
        std::vector<u64> t;
        for (u32 t0 = (u32)log_p & 3; t0 + 4 < (u32)log_p; t0 += 4) {
            const u32 log_js = (u32)log_p - t0 - 4;
            for (u32 r = 1; r < 16; r++) {
                const u64 step = h_gl_pow(w, (u64)ntt_brev4((int)r) << t0);       // w_P^(bitrev4(r) << t0)
                u64 v = 1;
                for (u32 lo = 0; lo < (1u << log_js); lo++) { t.push_back(v); v = h_gl_mul(v, step); }
            }
        }
        if (t.empty()) t.push_back(1);         // networks without a non-final radix-16 round: a valid (unused) pointer
        const u64 *d = upload(t);
        if (!d) return nullptr;
        return tau_cache[key] = d;
    }
    // two-level powers of w_n (inverse: w_n^-1)
    W2 w2(int log_n, bool inverse) {
        auto key = std::make_pair(log_n, (int)inverse);
        auto it = w2_cache.find(key);
        if (it != w2_cache.end()) return it->second;
        u32 lo_bits = (log_n + 1) / 2;
        size_t nlo = (size_t)1 << lo_bits, nhi = (size_t)1 << (log_n - lo_bits);
        std::vector<u64> lo(nlo), hi(nhi);
        u64 w = h_gl_root_of_unity(log_n);
        if (inverse) w = h_gl_inv(w);
        lo[0] = 1;
        for (size_t i = 1; i < nlo; i++) lo[i] = h_gl_mul(lo[i - 1], w);
        u64 wh = h_gl_pow(w, nlo);
        hi[0] = 1;
        for (size_t i = 1; i < nhi; i++) hi[i] = h_gl_mul(hi[i - 1], wh);
        W2 r = {upload(lo), upload(hi), lo_bits};
        if (!r.lo || !r.hi) return r;
        return w2_cache[key] = r;
    }
    // s_e = 7 * w_L^e, e < 2^rate_bits:  a[e][j] = s_e^(j * st), j < P;  b[e][r] = s_e^r, r < st;  st = n / P
    Shift shift(int log_n, int rate_bits, int log_p) {
        auto key = std::make_tuple(log_n, rate_bits, log_p);
        auto it = shift_cache.find(key);
        if (it != shift_cache.end()) return it->second;
        size_t P = (size_t)1 << log_p, st = (size_t)1 << (log_n - log_p), E = (size_t)1 << rate_bits;
        std::vector<u64> a(E * P), b(E * st);
        u64 wl = h_gl_root_of_unity(log_n + rate_bits);
        for (size_t e = 0; e < E; e++) {
            u64 s = h_gl_mul(7, h_gl_pow(wl, e));
            u64 sst = h_gl_pow(s, st);
            a[e * P] = 1;
            for (size_t j = 1; j < P; j++) a[e * P + j] = h_gl_mul(a[e * P + j - 1], sst);
            b[e * st] = 1;
            for (size_t r = 1; r < st; r++) b[e * st + r] = h_gl_mul(b[e * st + r - 1], s);
        }
        Shift r = {upload(a), upload(b)};
        if (!r.a || !r.b) return r;
        return shift_cache[key] = r;
    }
};

#ifndef NTT_TILE_LOG
#define NTT_TILE_LOG 13      // log2 of the elements staged per tile (when the transform is shorter than that)
#endif
#ifndef NTT_MAX_THREADS
#define NTT_MAX_THREADS 512
#endif
static inline uint32_t ntt_threads(u32 log_p, u32 log_a) {
    u32 blocks = (1u << (log_p + log_a)) >> 4;  // radix-16 register blocks per tile
    u32 t = blocks < 32 ? 32 : blocks;
    return t > NTT_MAX_THREADS ? NTT_MAX_THREADS : t;
}
// lanes per tile: ~8K elements per tile (16K once P >= 2^11), never fewer than 4 lanes (32-byte segments) below 2^13
static inline u32 ntt_log_a_strided(u32 log_p, u32 log_st) {
    u32 la = log_p >= NTT_TILE_LOG ? 0 : NTT_TILE_LOG - log_p;
    const u32 la_min = log_p >= 13 ? 1 : 2;   // 2^13-point tiles only fit two lanes in shared memory
    if (la < la_min) la = la_min;
    if (la > 6) la = 6;
    if (la > log_st) la = log_st;
    return la;
}
static inline u32 ntt_log_a_contig(u32 log_p) {
    u32 la = log_p >= NTT_TILE_LOG ? 0 : NTT_TILE_LOG - log_p;
    if (la > 8) la = 8;
    return la;
}
static inline NttLaunch ntt_make_launch(int mode, const NttPass &p) {
    NttLaunch l;
    l.mode = mode; l.p = p;
    l.p.tau_in_smem = ntt_tau_fits(p.log_p, p.log_a) ? 1 : 0;
    l.threads = ntt_threads(p.log_p, p.log_a);
    l.smem = ntt_smem_bytes(p.log_p, p.log_a);
    return l;
}
static inline u32 ntt_split_first(u32 log_n) { return (log_n + 1) / 2; }

// values [C][n] (stride in_stride) -> coeffs [C][n] natural order.  scratch: [C][n], distinct from out.
// in may alias out (pass 1 consumes `in` completely before pass 2 writes `out`).  Returns false if unsupported.
static inline bool ntt_plan_intt(NttTableStore &ts, const u64 *in, u64 in_stride, u64 *scratch, u64 scratch_stride,
                                 u64 *out, u64 out_stride, u32 C, u32 log_n, std::vector<NttLaunch> &plan) {
    if (log_n > 2 * NTT_MAX_LOGP) return false;
    u64 ninv = h_gl_inv(((u64)1 << log_n) % GL_P);
    NttPass p = {};
    p.log_n = log_n; p.num_cols = C; p.scale = ninv; p.rate_bits = 0;
    if (log_n <= NTT_MAX_LOGP) {
        p.in = in; p.out = out; p.in_col_stride = in_stride; p.out_col_stride = out_stride;
        p.log_p = log_n; p.log_a = ntt_log_a_contig(log_n);
        p.num_tiles = ((u64)C + (1u << p.log_a) - 1) >> p.log_a;
        p.tw_local = ts.tw_local(log_n, true);

        p.tau_tab = ts.tau_tab(log_n, true); p.tau_in_smem = 0;
        plan.push_back(ntt_make_launch(NTT_INTT_SINGLE, p));
        return true;
    }
    auto w = ts.w2(log_n, true);
    u32 l2 = ntt_split_first(log_n), l1 = log_n - l2;  // N2 = 2^l2 points in pass 1, N1 = 2^l1 in pass 2
    p.w_lo = w.lo; p.w_hi = w.hi; p.w_lo_bits = w.lo_bits;
    // pass 1: tiles over n1
    p.in = in; p.in_col_stride = in_stride; p.out = scratch; p.out_col_stride = scratch_stride;
    p.log_p = l2; p.log_a = ntt_log_a_strided(l2, l1);
    p.num_tiles = (u64)C << (l1 - p.log_a);
    p.tw_local = ts.tw_local(l2, true);

    p.tau_tab = ts.tau_tab(l2, true); p.tau_in_smem = 0;
    plan.push_back(ntt_make_launch(NTT_INTT_P1, p));
    // pass 2: tiles over k2
    p.in = scratch; p.in_col_stride = scratch_stride; p.out = out; p.out_col_stride = out_stride;
    p.log_p = l1; p.log_a = ntt_log_a_strided(l1, l2);
    p.num_tiles = (u64)C << (l2 - p.log_a);
    p.tw_local = ts.tw_local(l1, true);

    p.tau_tab = ts.tau_tab(l1, true); p.tau_in_smem = 0;
    plan.push_back(ntt_make_launch(NTT_INTT_P2, p));
    return true;
}

// coeffs [C][n] -> lde (column-major, bit-reversed rows), shift 7, as 2^log_shards row shards: [G][C][L/G]
// (log_shards = 0: the plain [C][L] layout).  Returns false if the shape is unsupported.
// shard_out (optional, 2^log_shards pointers): the last pass stores row shard g at shard_out[g] ([C][L/G], possibly
// peer memory) instead of lde + g * C * L/G; lde is then only the local intermediate of the first pass.
// first_shard: with shard_out, the last pass starts at that row shard (the caller's own rank) and wraps around.
static inline bool ntt_plan_lde(NttTableStore &ts, const u64 *coeffs, u64 coeffs_stride, u64 *lde, u32 C, u32 log_n,
                                u32 rate_bits, u32 log_shards, std::vector<NttLaunch> &plan, u64 *const *shard_out = nullptr,
                                u32 first_shard = 0) {
    if (log_n > 2 * NTT_MAX_LOGP) return false;
    const u32 log_l = log_n + rate_bits;
    if (log_shards > log_l) return false;
    NttPass p = {};
    p.log_n = log_n; p.num_cols = C; p.rate_bits = rate_bits;
    p.log_shard_rows = log_l - log_shards;
    p.out_col_stride = (u64)1 << p.log_shard_rows;
    p.shard_stride = (u64)C << p.log_shard_rows;
    p.in = coeffs; p.in_col_stride = coeffs_stride; p.out = lde;
    if (shard_out && log_shards > 4) return false;
    auto set_shard_ptrs = [&](NttPass &q) {
        q.num_shard_ptrs = 0;
        if (!shard_out) return;
        q.num_shard_ptrs = 1u << log_shards;
        for (u32 g = 0; g < q.num_shard_ptrs; g++) q.shard_out[g] = shard_out[g];
    };
    if (log_n <= NTT_MAX_LOGP) {
        auto s = ts.shift(log_n, rate_bits, log_n);  // st = 1: a[e][j] = s_e^j is the whole table
        p.shift_a = s.a; p.shift_b = s.a;
        p.log_p = log_n; p.log_a = ntt_log_a_contig(log_n);
        u64 units = (u64)C << rate_bits;
        p.num_tiles = (units + (1u << p.log_a) - 1) >> p.log_a;
        p.tw_local = ts.tw_local(log_n, false);

        p.tau_tab = ts.tau_tab(log_n, false); p.tau_in_smem = 0;
        set_shard_ptrs(p);
        plan.push_back(ntt_make_launch(NTT_LDE_SINGLE, p));
        return true;
    }
    u32 lp = ntt_split_first(log_n), lst = log_n - lp;
    if (lst > p.log_shard_rows) return false;  // a contiguous last-pass run must not straddle row shards
    auto w = ts.w2(log_n, false);
    auto s = ts.shift(log_n, rate_bits, lp);
    p.w_lo = w.lo; p.w_hi = w.hi; p.w_lo_bits = w.lo_bits; p.shift_a = s.a; p.shift_b = s.b;
    p.log_p = lp; p.log_a = ntt_log_a_strided(lp, lst);
    p.num_tiles = ((u64)C << rate_bits) << (lst - p.log_a);
    p.tw_local = ts.tw_local(lp, false);

    p.tau_tab = ts.tau_tab(lp, false); p.tau_in_smem = 0;
    plan.push_back(ntt_make_launch(NTT_LDE_FIRST, p));
    // last pass: contiguous runs of 2^lst points, in place over the whole LDE buffer
    p.in = lde;
    p.log_p = lst; p.log_a = ntt_log_a_contig(lst);
    u64 units = (u64)C << (log_n + rate_bits - lst);
    p.num_units = units;
    p.num_tiles = (units + (1u << p.log_a) - 1) >> p.log_a;
    p.tw_local = ts.tw_local(lst, false);

    p.tau_tab = ts.tau_tab(lst, false); p.tau_in_smem = 0;
    set_shard_ptrs(p);
    if (shard_out && (p.num_tiles % ((u64)1 << log_shards)) == 0)
        p.tile_rot = (p.num_tiles >> log_shards) * (first_shard & ((1u << log_shards) - 1));
    plan.push_back(ntt_make_launch(NTT_DIF_LAST, p));
    return true;
}
