// libplonky2_b200.so -- engine context, device-resident PolynomialBatch handles and the C ABI of
// include/plonky2_b200.h.  Host orchestration only; the arithmetic lives in ntt.cuh / merkle.cuh / poseidon.cuh.
//
// Replaces, for the hot path, plonky2::fri::oracle::PolynomialBatch::{from_values, from_coeffs, get_lde_values} and
// plonky2::hash::merkle_tree::MerkleTree::{new, get, prove} (dep plonky2 0.1.4, /root/reference/Cargo.lock:2347-2350),
// reached from /root/reference/eth-lc-plonky2/src/main.rs:227 (build) and :230 (prove).
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/plonky2_b200.h"
#include "merkle.cuh"
#include "microbench.cuh"
#include "ntt.cuh"
#include "ntt_plan.h"
#include "prover.cuh"
#include "plonk.cuh"
#include <map>
#include <sys/random.h>

#define SALT_SIZE 4u

namespace {

thread_local std::string g_err;
std::recursive_mutex g_mu;

eng_status fail(eng_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? ENG_ERR_OOM : ENG_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)
#define ST(call)                         \
    do {                                 \
        eng_status s_ = (call);          \
        if (s_ != ENG_OK) return s_;     \
    } while (0)

struct Ctx {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    NttTableStore tables;
    std::vector<void *> table_allocs;
    uint64_t launches = 0;
    cudaEvent_t ev[8] = {};
    std::map<std::pair<u64, u32>, NttTableStore::W2> pow_cache;   // two-level power tables of arbitrary bases
    std::multimap<size_t, void *> free_blocks;      // exact-size cache of released device buffers (dev_alloc / dev_free)
    std::map<void *, size_t> live_blocks;
    size_t cached_bytes = 0, cache_cap = (size_t)64 << 30;
    // host-input pipeline: copies run on copy_stream while the previous column chunk transforms on `stream`
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_ev = nullptr, fork_ev = nullptr;
    u64 *stage[2] = {nullptr, nullptr};      // pinned bounce buffers for pageable sources
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    size_t stage_elems = 0;
    int stage_next = 0;
    bool table_upload_failed = false;        // a twiddle / power table could not be placed on the device (reported by tables_ok)
    // eng_set_option
    int opt_native_gates = 1;                // quotient: library gates through the compiled evaluators (0: through their bytecode)
    int opt_native_poseidon = 1;             // quotient: PoseidonGate through the native FP64 evaluator (0: its bytecode)
    int opt_peer_chunk_cols = 4;             // eng_lde_peer_dev: columns per iNTT + LDE chunk (0: the whole column shard at once)
    int opt_reserve = 1;                     // eng_circuit_new / eng_circuit_load grow the pool to one proof's footprint
    int opt_lde_group_mb = 0;                // LDE: megabytes of one (column, coset) group kept between the two passes (L2 residency)
};
Ctx g;

}  // namespace

struct eng_batch {
    uint32_t num_polys = 0, degree_log = 0, rate_bits = 0, cap_height = 0, blinding = 0, leaf_len = 0;
    uint64_t num_leaves = 0, num_digests = 0;
    uint32_t num_layers = 0;
    u64 *coeffs = nullptr;       // [num_polys][n]
    u64 *lde = nullptr;          // [leaf_len][num_leaves] column-major, bit-reversed rows (owned)
    const u64 *leaf_data = nullptr;  // what the tree was built over
    u64 row_stride = 0, col_stride = 0;
    u64 *digests = nullptr, *cap = nullptr;
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
};

namespace {

// Blinding salts (PolynomialBatch::from_coeffs with blinding: SALT_SIZE random elements per leaf; plonky2 draws them from
// OsRng through F::rand_vec).  Here: the ChaCha20 key stream under a 256-bit key -- taken from the OS RNG (getrandom) for
// every batch unless the caller asks for the reproducible test stream -- block counter = thread index.  One 64-byte block
// gives 4 salts from its first four 64-bit words; a word >= p (probability 2^-32) is replaced by the next spare word of the
// block (rejection sampling, as rand's gen_range does), so the salts are uniform in [0, p) and one opened leaf says
// nothing about the others.
struct SaltKey { u32 k[8]; };
__device__ __forceinline__ u32 salt_rotl(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
#define SALT_QR(a, b, c, d) a += b; d = salt_rotl(d ^ a, 16); c += d; b = salt_rotl(b ^ c, 12); a += b; d = salt_rotl(d ^ a, 8); c += d; b = salt_rotl(b ^ c, 7);
__global__ void salt_kernel(u64 *dst, u64 count, SaltKey key) {
    const u64 blk = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (blk * 4 >= count) return;
    u32 st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                  key.k[4], key.k[5], key.k[6], key.k[7], (u32)blk, (u32)(blk >> 32), 0u, 0u};
    u32 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = st[i];
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        SALT_QR(x[0], x[4], x[8], x[12]) SALT_QR(x[1], x[5], x[9], x[13]) SALT_QR(x[2], x[6], x[10], x[14]) SALT_QR(x[3], x[7], x[11], x[15])
        SALT_QR(x[0], x[5], x[10], x[15]) SALT_QR(x[1], x[6], x[11], x[12]) SALT_QR(x[2], x[7], x[8], x[13]) SALT_QR(x[3], x[4], x[9], x[14])
    }
    u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = (u64)(x[2 * i] + st[2 * i]) | ((u64)(x[2 * i + 1] + st[2 * i + 1]) << 32);
    int spare = 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 v = w[i];
        while (v >= GL_P && spare < 8) v = w[spare++];
        if (blk * 4 + i < count) dst[blk * 4 + i] = gl_canon(v);   // all spares used (2^-160): canonicalise
    }
}
// key of one batch: 32 bytes of OS randomness, or -- ONLY for reproducible tests -- the ChaCha-style expansion of a seed
eng_status salt_key(uint64_t seed, SaltKey &key) {
    if (seed == 0) {
        size_t got = 0;
        while (got < sizeof(key.k)) {
            ssize_t r = getrandom((char *)key.k + got, sizeof(key.k) - got, 0);
            if (r <= 0) return fail(ENG_ERR_STATE, "getrandom failed: blinding needs the OS random number generator");
            got += (size_t)r;
        }
        return ENG_OK;
    }
    u64 z = seed;
    for (int i = 0; i < 4; i++) {
        z += 0x9E3779B97F4A7C15ULL;
        u64 v = z;
        v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ULL;
        v = (v ^ (v >> 27)) * 0x94D049BB133111EBULL;
        v ^= v >> 31;
        key.k[2 * i] = (u32)v; key.k[2 * i + 1] = (u32)(v >> 32);
    }
    return ENG_OK;
}
__global__ void permute_kernel(const u64 *in, u64 *out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
    poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[12 * i + k] = gl_canon(s[k]);
}
__global__ void two_to_one_kernel(const u64 *pairs, u64 *out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 d[4];
    poseidon_two_to_one(pairs + 8 * i, pairs + 8 * i + 4, d);
#pragma unroll
    for (int k = 0; k < 4; k++) out[4 * i + k] = d[k];
}
// rows [first, first+count) of the leaf matrix -> row-major out[count][width]
__global__ void gather_rows_kernel(const u64 *data, u64 row_stride, u64 col_stride, u32 width, u64 first, u64 count,
                                   u64 *out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * width) return;
    u64 r = i / width, c = i % width;
    out[i] = gl_canon(data[(first + r) * row_stride + c * col_stride]);
}

// Device buffers: an exact-size cache in front of the stream-ordered pool.  A proof (and a bench step) repeats the same
// sequence of multi-GB requests; handing freed blocks back by exact size makes every request after the first round a
// cache hit.  Going through cudaMallocAsync / cudaFreeAsync each time let the pool split and coalesce its blocks
// differently from call to call, and every so often a 9 GB request went back to the driver (measured: sporadic
// 0.2 ... 1 s stages in a 0.25 s proof).  All engine work is ordered on g.stream, so reuse needs no extra events.
// ENG_TRACE=1: host wall-clock of the slow first-call steps (pool growth, table uploads) on stderr
bool trace_on() { static int v = -1; if (v < 0) { const char *e = getenv("ENG_TRACE"); v = (e && *e && *e != '0') ? 1 : 0; } return v == 1; }
double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// trace_mark("label"): with ENG_TRACE set, waits for the device and prints the wall clock since the previous mark
void trace_mark(const char *label) {
    if (!trace_on()) return;
    static double last = 0.0;
    cudaDeviceSynchronize();
    const double now = wall_ms();
    fprintf(stderr, "[eng trace] %-44s +%9.1f ms\n", label, last == 0.0 ? 0.0 : now - last);
    last = now;
}
void cache_flush() {
    for (auto &kv : g.free_blocks) cudaFreeAsync(kv.second, g.stream);
    g.free_blocks.clear();
    g.cached_bytes = 0;
}
eng_status dev_alloc(u64 **p, size_t elems) {
    *p = nullptr;
    if (elems == 0) return ENG_OK;
    const size_t bytes = elems * sizeof(u64);
    auto it = g.free_blocks.find(bytes);
    if (it != g.free_blocks.end()) {
        *p = (u64 *)it->second;
        g.free_blocks.erase(it);
        g.cached_bytes -= bytes;
        g.live_blocks[*p] = bytes;
        return ENG_OK;
    }
    const double t0 = trace_on() ? wall_ms() : 0.0;
    cudaError_t e = cudaMallocAsync((void **)p, bytes, g.stream);
    if (trace_on() && bytes >= ((size_t)64 << 20)) {
        cudaStreamSynchronize(g.stream);
        fprintf(stderr, "[eng trace] pool miss: %.2f GB in %.1f ms\n", bytes / 1e9, wall_ms() - t0);
    }
    if (e != cudaSuccess && !g.free_blocks.empty()) {   // give the cached blocks back and retry once
        cudaGetLastError();
        cache_flush();
        cudaStreamSynchronize(g.stream);
        e = cudaMallocAsync((void **)p, bytes, g.stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(ENG_ERR_OOM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    g.live_blocks[*p] = bytes;
    return ENG_OK;
}
void dev_free(void *p) {
    if (!p) return;
    auto it = g.live_blocks.find(p);
    if (it == g.live_blocks.end()) { cudaFreeAsync(p, g.stream); return; }
    const size_t bytes = it->second;
    g.live_blocks.erase(it);
    if (g.cached_bytes + bytes > g.cache_cap) cache_flush();   // many different shapes: do not hoard the device
    g.free_blocks.emplace(bytes, p);
    g.cached_bytes += bytes;
}

// Scope guard for temporary device buffers: every early return releases them (error paths leaked).
struct DevBuf {
    u64 *p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    eng_status alloc(size_t elems) { release(); return dev_alloc(&p, elems); }
    void release() { if (p) { dev_free(p); p = nullptr; } }
    u64 *take() { u64 *q = p; p = nullptr; return q; }
};

// A table upload that failed is not cached (NttTableStore drops it) and poisons nothing: the call that needed it fails here.
eng_status tables_ok() {
    if (!g.table_upload_failed) return ENG_OK;
    g.table_upload_failed = false;
    return fail(ENG_ERR_OOM, "device allocation of a twiddle table failed");
}

template <int MODE>
eng_status launch_mode(const NttLaunch &l) {
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(ntt_pass_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CU(cudaFuncSetAttribute(ntt_pass_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ntt_pass_kernel<MODE>, (int)l.threads, l.smem));
    if (per_sm < 1) return fail(ENG_ERR_CUDA, "NTT pass (log_p=%u log_a=%u) does not fit on an SM", l.p.log_p, l.p.log_a);
    u64 grid = (u64)g.sm_count * per_sm;
    if (grid > l.p.num_tiles) grid = l.p.num_tiles;
    if (grid == 0) return ENG_OK;
    ntt_pass_kernel<MODE><<<(unsigned)grid, l.threads, l.smem, g.stream>>>(l.p);
    g.launches++;
    CU(cudaGetLastError());
    return ENG_OK;
}
eng_status launch_plan(const std::vector<NttLaunch> &plan) {
    ST(tables_ok());
    for (const NttLaunch &l : plan) {
        switch (l.mode) {
            case NTT_LDE_FIRST: ST(launch_mode<NTT_LDE_FIRST>(l)); break;
            case NTT_LDE_SINGLE: ST(launch_mode<NTT_LDE_SINGLE>(l)); break;
            case NTT_DIF_LAST: ST(launch_mode<NTT_DIF_LAST>(l)); break;
            case NTT_INTT_P1: ST(launch_mode<NTT_INTT_P1>(l)); break;
            case NTT_INTT_P2: ST(launch_mode<NTT_INTT_P2>(l)); break;
            case NTT_INTT_SINGLE: ST(launch_mode<NTT_INTT_SINGLE>(l)); break;
            default: return fail(ENG_ERR_INVALID, "unknown NTT pass mode %d", l.mode);
        }
    }
    return ENG_OK;
}

// MerkleTree::new over b->leaf_data; fills digests and cap.
eng_status build_tree(eng_batch *b) {
    MerkleParams mp;
    mp.data = b->leaf_data; mp.row_stride = b->row_stride; mp.col_stride = b->col_stride;
    mp.width = b->leaf_len; mp.num_leaves = b->num_leaves; mp.num_layers = b->num_layers;
    mp.digests = b->digests; mp.cap = b->cap;
    mp.noop_max = 4;  // H::hash_or_noop
    CU(cudaEventRecord(g.ev[3], g.stream));
    {
        unsigned threads = 128;
        u64 blocks = (b->num_leaves + threads - 1) / threads;
        merkle_leaves_kernel<<<(unsigned)blocks, threads, 0, g.stream>>>(mp);
        g.launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(g.ev[4], g.stream));
    for (u32 layer = 0; layer < b->num_layers; layer++) {
        u64 parents = b->num_leaves >> (layer + 1);
        unsigned threads = 128;
        u64 blocks = (parents + threads - 1) / threads;
        merkle_level_kernel<<<(unsigned)blocks, threads, 0, g.stream>>>(mp, layer);
        g.launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(g.ev[5], g.stream));
    return ENG_OK;
}

eng_status check_ready() {
    if (!g.ready) return fail(ENG_ERR_STATE, "engine not initialised: call eng_init(device) on a machine with a CUDA device");
    return ENG_OK;
}

eng_status collect_times(eng_batch *b, bool had_ifft, bool had_lde) {
    CU(cudaStreamSynchronize(g.stream));
    float ms = 0;
    if (had_ifft) { CU(cudaEventElapsedTime(&ms, g.ev[1], g.ev[2])); b->stage_ms[0] = ms; }
    if (had_lde) { CU(cudaEventElapsedTime(&ms, g.ev[2], g.ev[3])); b->stage_ms[1] = ms; }
    CU(cudaEventElapsedTime(&ms, g.ev[3], g.ev[4])); b->stage_ms[3] = ms;
    CU(cudaEventElapsedTime(&ms, g.ev[4], g.ev[5])); b->stage_ms[4] = ms;
    if (had_ifft || had_lde) { CU(cudaEventElapsedTime(&ms, g.ev[0], g.ev[1])); b->stage_ms[5] = ms; }
    return ENG_OK;
}

void destroy_batch(eng_batch *b) {
    if (!b) return;
    dev_free(b->coeffs); dev_free(b->lde); dev_free(b->digests); dev_free(b->cap);
    delete b;
}


// ---- host -> device column copies -------------------------------------------------------------------------------------
// A PolynomialBatch arrives as C host columns (plonky2: Vec<PolynomialValues<F>>).  Page-locked (or registered) columns go
// straight to the copy engine.  Pageable columns -- what a Rust Vec is -- would make cudaMemcpyAsync a single-threaded
// bounce copy (~10 GB/s), so they are staged through two pinned buffers by a few host threads and sent from there.
constexpr size_t STAGE_BYTES = (size_t)64 << 20;
// test hooks: ENG_H2D_STAGE_BYTES / ENG_H2D_CHUNK_BYTES shrink the bounce buffers / the column chunk so that small
// parity cases run through several chunks and several bounce-buffer refills
size_t env_bytes(const char *name, size_t dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    unsigned long long x = strtoull(v, nullptr, 10);
    return x >= 64 ? (size_t)x : dflt;
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

eng_status ensure_stage() {
    const size_t bytes = env_bytes("ENG_H2D_STAGE_BYTES", STAGE_BYTES);
    if (g.stage[0] && g.stage_elems == bytes / sizeof(u64)) return ENG_OK;
    for (int i = 0; i < 2; i++) {
        if (g.stage[i]) { CU(cudaEventSynchronize(g.stage_ev[i])); CU(cudaFreeHost(g.stage[i])); g.stage[i] = nullptr; }
        if (!g.stage_ev[i]) CU(cudaEventCreateWithFlags(&g.stage_ev[i], cudaEventDisableTiming));
        CU(cudaHostAlloc((void **)&g.stage[i], bytes, cudaHostAllocDefault));
    }
    g.stage_elems = bytes / sizeof(u64);
    return ENG_OK;
}

// dst[cc][n] (device, contiguous) <- cols[0..cc) (host), enqueued on g.copy_stream.
eng_status h2d_columns(const uint64_t *const *cols, uint32_t cc, u64 n, u64 *dst) {
    bool pinned = true;
    for (uint32_t c = 0; c < cc && pinned; c++) pinned = host_ptr_is_pinned(cols[c]);
    if (pinned) {
        for (uint32_t c = 0; c < cc; c++)
            CU(cudaMemcpyAsync(dst + (size_t)c * n, cols[c], n * sizeof(u64), cudaMemcpyHostToDevice, g.copy_stream));
        return ENG_OK;
    }
    ST(ensure_stage());
    const u64 total = (u64)cc * n;
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned T = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 1);
    for (u64 s0 = 0; s0 < total; s0 += g.stage_elems) {
        const u64 len = total - s0 < g.stage_elems ? total - s0 : g.stage_elems;
        const int slot = g.stage_next;
        g.stage_next ^= 1;
        CU(cudaEventSynchronize(g.stage_ev[slot]));      // the previous copy out of this buffer has finished
        u64 *buf = g.stage[slot];
        auto work = [&](unsigned t) {                     // flat range [a, b) of the virtual concatenation of the columns
            u64 a = s0 + len * t / T, b = s0 + len * (t + 1) / T;
            while (a < b) {
                u64 c = a / n, off = a % n;
                u64 run = n - off < b - a ? n - off : b - a;
                memcpy(buf + (a - s0), cols[c] + off, run * sizeof(u64));
                a += run;
            }
        };
        if (T == 1 || len < (1u << 16)) {
            for (unsigned t = 0; t < T; t++) work(t);
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < T; t++) th.emplace_back(work, t);
            work(0);
            for (auto &x : th) x.join();
        }
        CU(cudaMemcpyAsync(dst + s0, buf, len * sizeof(u64), cudaMemcpyHostToDevice, g.copy_stream));
        CU(cudaEventRecord(g.stage_ev[slot], g.copy_stream));
    }
    return ENG_OK;
}

// Shared tail of from_values / from_coeffs.  `src` holds values (is_values) or coefficients, either as C host
// column pointers (cols_host) or as one device array [C][n] (src_dev).
// values_keep_dev (optional, host input only): receives a device copy [C][n] of the input columns as they arrive.
eng_status make_batch(const uint64_t *const *cols_host, const u64 *src_dev, bool is_values, uint32_t C, uint32_t log_n,
                      uint32_t rate_bits, int32_t blinding, uint64_t seed, uint32_t cap_height, eng_batch **out,
                      u64 *values_keep_dev = nullptr) {
    ST(check_ready());
    if (!out) return fail(ENG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (C == 0) return fail(ENG_ERR_INVALID, "PolynomialBatch needs at least one polynomial");
    if (log_n > 2 * NTT_MAX_LOGP) return fail(ENG_ERR_INVALID, "degree_log %u > %d unsupported", log_n, 2 * NTT_MAX_LOGP);
    if (log_n + rate_bits > 32) return fail(ENG_ERR_INVALID, "LDE size 2^%u exceeds the field's two-adicity", log_n + rate_bits);
    if (cap_height > log_n + rate_bits)
        return fail(ENG_ERR_INVALID, "cap_height %u > log2(leaves) %u (MerkleTree::new assertion)", cap_height, log_n + rate_bits);
    if (!cols_host && !src_dev) return fail(ENG_ERR_INVALID, "input is NULL");

    eng_batch *b = new eng_batch();
    b->num_polys = C; b->degree_log = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
    b->blinding = blinding ? 1 : 0;
    b->leaf_len = C + (blinding ? SALT_SIZE : 0);
    const u64 n = (u64)1 << log_n, L = n << rate_bits;
    b->num_leaves = L;
    b->num_layers = log_n + rate_bits - cap_height;
    b->num_digests = 2 * (L - ((u64)1 << cap_height));
    eng_status st;
    if ((st = dev_alloc(&b->coeffs, (size_t)C * n)) != ENG_OK || (st = dev_alloc(&b->lde, (size_t)b->leaf_len * L)) != ENG_OK ||
        (st = dev_alloc(&b->digests, (size_t)b->num_digests * 4)) != ENG_OK ||
        (st = dev_alloc(&b->cap, (size_t)4 << cap_height)) != ENG_OK) {
        destroy_batch(b);
        return st;
    }
    b->leaf_data = b->lde; b->row_stride = 1; b->col_stride = L;
    trace_mark("make_batch: buffers allocated");

    auto bail = [&](eng_status s) { if (g.copy_stream) cudaStreamSynchronize(g.copy_stream); destroy_batch(b); return s; };
#define STB(call) do { eng_status s__ = (call); if (s__ != ENG_OK) return bail(s__); } while (0)
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(ENG_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__))); } while (0)

    CUB(cudaEventRecord(g.ev[0], g.stream));
    std::vector<NttLaunch> plan;
    if (cols_host) {
        // Host columns: pipelined by column chunk.  Chunk k is copied on the copy stream while chunk k-1 runs its iNTT
        // and coset LDE on the compute stream, so the transfer hides behind the transforms (the leaf hash needs every
        // column and follows).  The iNTT scratch of a chunk is the LDE region of its own columns, not yet written.
        for (uint32_t c = 0; c < C; c++)
            if (!cols_host[c]) return bail(fail(ENG_ERR_INVALID, "column %u is NULL", c));
        u64 per = env_bytes("ENG_H2D_CHUNK_BYTES", (size_t)32 << 20) / (n * sizeof(u64));
        const uint32_t chunk = (uint32_t)(per < 1 ? 1 : (per > C ? C : per));
        CUB(cudaEventRecord(g.fork_ev, g.stream));                 // the buffers were allocated in g.stream's order
        CUB(cudaStreamWaitEvent(g.copy_stream, g.fork_ev, 0));
        for (uint32_t c0 = 0; c0 < C; c0 += chunk) {
            const uint32_t cc = C - c0 < chunk ? C - c0 : chunk;
            u64 *co = b->coeffs + (size_t)c0 * n, *ld = b->lde + (size_t)c0 * L;
            STB(h2d_columns(cols_host + c0, cc, n, co));
            CUB(cudaEventRecord(g.chunk_ev, g.copy_stream));
            CUB(cudaStreamWaitEvent(g.stream, g.chunk_ev, 0));
            if (values_keep_dev)
                CUB(cudaMemcpyAsync(values_keep_dev + (size_t)c0 * n, co, (size_t)cc * n * sizeof(u64), cudaMemcpyDeviceToDevice, g.stream));
            if (is_values) {
                plan.clear();
                if (!ntt_plan_intt(g.tables, co, n, ld, n, co, n, cc, log_n, plan)) return bail(fail(ENG_ERR_INVALID, "iNTT size unsupported"));
                STB(launch_plan(plan));
            }
            plan.clear();
            if (!ntt_plan_lde(g.tables, co, n, ld, cc, log_n, rate_bits, 0, plan)) return bail(fail(ENG_ERR_INVALID, "LDE size unsupported"));
            STB(launch_plan(plan));
        }
        CUB(cudaEventRecord(g.ev[1], g.stream));   // stage_ms: "host to device" = the whole pipelined copy + transforms
        CUB(cudaEventRecord(g.ev[2], g.stream));
    } else {
        const u64 *src = src_dev;
        CUB(cudaEventRecord(g.ev[1], g.stream));
        // Column groups (option lde_group_mb): the two passes of a transform run group by group, so that what pass 1 wrote
        // (the four-step intermediate, one LDE column = 8 * L bytes) is still in the 126 MB L2 when pass 2 reads it --
        // the intermediate then makes no HBM round trip.  0: one launch per pass over the whole batch.
        const u64 group_bytes = (u64)g.opt_lde_group_mb << 20;
        const uint32_t cg_lde = group_bytes ? (uint32_t)std::min<u64>(C, std::max<u64>(1, group_bytes / (L * sizeof(u64)))) : C;
        const uint32_t cg_intt = group_bytes ? (uint32_t)std::min<u64>(C, std::max<u64>(1, group_bytes / (n * sizeof(u64)))) : C;
        if (is_values) {
            // iNTT; the (not yet written) LDE buffer is the four-step scratch
            for (uint32_t c0 = 0; c0 < C; c0 += cg_intt) {
                const uint32_t cc = std::min(cg_intt, C - c0);
                plan.clear();
                if (!ntt_plan_intt(g.tables, src + (size_t)c0 * n, n, b->lde + (size_t)c0 * n, n, b->coeffs + (size_t)c0 * n, n, cc, log_n, plan))
                    return bail(fail(ENG_ERR_INVALID, "iNTT size unsupported"));
                STB(launch_plan(plan));
            }
        } else if (src != b->coeffs) {
            CUB(cudaMemcpyAsync(b->coeffs, src, (size_t)C * n * sizeof(u64), cudaMemcpyDeviceToDevice, g.stream));
        }
        CUB(cudaEventRecord(g.ev[2], g.stream));
        for (uint32_t c0 = 0; c0 < C; c0 += cg_lde) {
            const uint32_t cc = std::min(cg_lde, C - c0);
            plan.clear();
            if (!ntt_plan_lde(g.tables, b->coeffs + (size_t)c0 * n, n, b->lde + (size_t)c0 * L, cc, log_n, rate_bits, 0, plan))
                return bail(fail(ENG_ERR_INVALID, "LDE size unsupported"));
            STB(launch_plan(plan));
        }
    }
    if (blinding) {
        u64 count = (u64)SALT_SIZE * L;
        SaltKey key;
        STB(salt_key(seed, key));
        salt_kernel<<<(unsigned)((count / 4 + 255) / 256), 256, 0, g.stream>>>(b->lde + (size_t)C * L, count, key);
        g.launches++;
        CUB(cudaGetLastError());
    }
    trace_mark("make_batch: copies + transforms done");
    STB(build_tree(b));
    STB(collect_times(b, is_values, true));
    trace_mark("make_batch: tree done");
#undef STB
#undef CUB
    *out = b;
    return ENG_OK;
}

eng_status d2h(void *dst, const void *src, size_t bytes) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return ENG_OK;
}

}  // namespace

extern "C" {

eng_status eng_init(int32_t device) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (g.ready) {
        if (device != g.device) return fail(ENG_ERR_STATE, "engine already initialised on device %d", g.device);
        return ENG_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(ENG_ERR_STATE, "no CUDA device: %s (this engine has no CPU path)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(ENG_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    const double t_init0 = trace_on() ? wall_ms() : 0.0;
    CU(cudaSetDevice(device));
    CU(cudaFree(nullptr));   // forces context creation here (so that the trace attributes it)
    if (trace_on()) fprintf(stderr, "[eng trace] eng_init: CUDA context %.1f ms\n", wall_ms() - t_init0);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    g.cache_cap = (size_t)(0.6 * (double)prop.totalGlobalMem);
    CU(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (auto &ev : g.ev) CU(cudaEventCreate(&ev));
    CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g.chunk_ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.fork_ev, cudaEventDisableTiming));
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    CU(poseidon_upload_constants());
    g.tables.upload = [](const std::vector<u64> &t) -> const u64 * {
        void *d = nullptr;
        if (cudaMalloc(&d, t.size() * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); g.table_upload_failed = true; return nullptr; }
        if (cudaMemcpy(d, t.data(), t.size() * sizeof(u64), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d); g.table_upload_failed = true; return nullptr;
        }
        g.table_allocs.push_back(d);
        return (const u64 *)d;
    };
    g.launches = 0;
    g.ready = true;
    if (trace_on()) fprintf(stderr, "[eng trace] eng_init: total %.1f ms\n", wall_ms() - t_init0);
    return ENG_OK;
}

eng_status eng_shutdown(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!g.ready) return ENG_OK;
    cudaSetDevice(g.device);
    cache_flush();
    g.live_blocks.clear();
    cudaDeviceSynchronize();
    for (void *p : g.table_allocs) cudaFree(p);
    g.table_allocs.clear();
    g.tables.tw_local_cache.clear(); g.tables.tau_cache.clear(); g.tables.w2_cache.clear(); g.tables.shift_cache.clear(); g.pow_cache.clear();
    for (auto &ev : g.ev) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    if (g.own_stream) cudaStreamDestroy(g.own_stream);
    if (g.copy_stream) cudaStreamDestroy(g.copy_stream);
    if (g.chunk_ev) cudaEventDestroy(g.chunk_ev);
    if (g.fork_ev) cudaEventDestroy(g.fork_ev);
    for (int i = 0; i < 2; i++) {
        if (g.stage[i]) cudaFreeHost(g.stage[i]);
        if (g.stage_ev[i]) cudaEventDestroy(g.stage_ev[i]);
        g.stage[i] = nullptr; g.stage_ev[i] = nullptr;
    }
    g.copy_stream = nullptr; g.chunk_ev = g.fork_ev = nullptr;
    g.own_stream = g.stream = nullptr;
    g.ready = false;
    return ENG_OK;
}

eng_status eng_last_error(char *buf, size_t len) {
    if (!buf || len == 0) return ENG_ERR_INVALID;
    snprintf(buf, len, "%s", g_err.c_str());
    return ENG_OK;
}

eng_status eng_set_stream(void *cuda_stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    CU(cudaStreamSynchronize(g.stream));
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own_stream;
    return ENG_OK;
}

eng_status eng_release_cached(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    cache_flush();
    CU(cudaStreamSynchronize(g.stream));
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, g.device));
    CU(cudaMemPoolTrimTo(pool, 0));
    return ENG_OK;
}

// Grows the stream-ordered pool to hold `bytes` of free memory now (one driver call instead of one per buffer of the
// first proof: the pool maps device memory at ~50 GB/s, which the first eng_prove of a process otherwise pays buffer
// by buffer).  Clamped to what the device has free; failure to grow is not an error (the pool grows on demand).
eng_status eng_reserve(size_t bytes) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    // blocks parked in the engine's exact-size cache count as "used" for the pool: hand them back first, so that they are
    // part of the idle memory the reservation is measured against (and usable for the new shapes)
    cache_flush();
    CU(cudaStreamSynchronize(g.stream));
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, g.device));
    uint64_t reserved = 0, used = 0;
    CU(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
    CU(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
    const uint64_t idle = reserved > used ? reserved - used : 0;     // includes the engine's exact-size cache
    if (bytes <= idle) return ENG_OK;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t margin = (size_t)2 << 30;
    size_t want = bytes - idle;
    if (free_b <= margin) return ENG_OK;
    if (want > free_b - margin) want = free_b - margin;
    const double t0 = trace_on() ? wall_ms() : 0.0;
    void *p = nullptr;
    if (cudaMallocAsync(&p, want, g.stream) != cudaSuccess) { cudaGetLastError(); return ENG_OK; }
    cudaFreeAsync(p, g.stream);
    if (trace_on()) { cudaStreamSynchronize(g.stream); fprintf(stderr, "[eng trace] eng_reserve: pool grown by %.2f GB in %.1f ms\n", want / 1e9, wall_ms() - t0); }
    return ENG_OK;
}

// Page-locks a caller-owned host range (a witness column block, a Rust Vec) so that the column copies of
// eng_batch_from_values / eng_prove / eng_lde_peer_host / eng_h2d_columns go straight to the copy engine at PCIe speed
// instead of through the pinned bounce buffers (pageable memory: ~10 GB/s per process and host threads per copy).
eng_status eng_host_register(const void *ptr, size_t bytes) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!ptr || !bytes) return fail(ENG_ERR_INVALID, "NULL / empty range");
    cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return ENG_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ENG_ERR_CUDA, "cudaHostRegister of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    return ENG_OK;
}
eng_status eng_host_unregister(const void *ptr) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!ptr) return ENG_OK;
    cudaError_t e = cudaHostUnregister(const_cast<void *>(ptr));
    if (e != cudaSuccess) { cudaGetLastError(); if (e != cudaErrorHostMemoryNotRegistered) return fail(ENG_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return ENG_OK;
}

eng_status eng_synchronize(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    CU(cudaStreamSynchronize(g.stream));
    return ENG_OK;
}

eng_status eng_set_option(const char *name, int64_t value) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!name) return fail(ENG_ERR_INVALID, "NULL option name");
    if (!strcmp(name, "quot_native_gates")) { g.opt_native_gates = value != 0; return ENG_OK; }
    if (!strcmp(name, "quot_native_poseidon")) { g.opt_native_poseidon = value != 0; return ENG_OK; }
    if (!strcmp(name, "lde_peer_chunk_cols")) { g.opt_peer_chunk_cols = value < 0 ? 0 : (int)value; return ENG_OK; }
    if (!strcmp(name, "reserve_for_proof")) { g.opt_reserve = value != 0; return ENG_OK; }
    if (!strcmp(name, "lde_group_mb")) { g.opt_lde_group_mb = value < 0 ? 0 : (int)value; return ENG_OK; }
    return fail(ENG_ERR_INVALID, "unknown option '%s'", name);
}

eng_status eng_launch_count(uint64_t *out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!out) return ENG_ERR_INVALID;
    *out = g.launches;
    return ENG_OK;
}

eng_status eng_measure_int_peak(double *ops_per_s) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!ops_per_s) return fail(ENG_ERR_INVALID, "NULL argument");
    uint32_t *d_out;
    CU(cudaMalloc((void **)&d_out, 64));
    const int iters = 4096, blocks = g.sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int kind = 0; kind < 3; kind++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, g.stream));
            if (kind == 0) int_peak_kernel<0><<<blocks, threads, 0, g.stream>>>(d_out, 12345u + rep, iters);
            if (kind == 1) int_peak_kernel<1><<<blocks, threads, 0, g.stream>>>(d_out, 12345u + rep, iters);
            if (kind == 2) int_peak_kernel<2><<<blocks, threads, 0, g.stream>>>(d_out, 12345u + rep, iters);
            g.launches++;
            CU(cudaEventRecord(e1, g.stream));
            CU(cudaEventSynchronize(e1));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        double ops = (double)blocks * threads * iters * 32.0 * (kind == 2 ? 2.0 : 1.0);
        ops_per_s[kind] = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return ENG_OK;
}

eng_status eng_poseidon_permute(const uint64_t *states_host, uint64_t *out_host, size_t count) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (count == 0) return ENG_OK;
    if (!states_host || !out_host) return fail(ENG_ERR_INVALID, "NULL buffer");
    DevBuf d_in, d_out;
    ST(d_in.alloc(count * 12));
    ST(d_out.alloc(count * 12));
    CU(cudaMemcpyAsync(d_in.p, states_host, count * 96, cudaMemcpyHostToDevice, g.stream));
    permute_kernel<<<(unsigned)((count + 127) / 128), 128, 0, g.stream>>>(d_in.p, d_out.p, count);
    g.launches++;
    CU(cudaGetLastError());
    return d2h(out_host, d_out.p, count * 96);
}

eng_status eng_hash_n(const uint64_t *in_host, size_t len, size_t count, int32_t or_noop, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (count == 0) return ENG_OK;
    if ((!in_host && len) || !out_host) return fail(ENG_ERR_INVALID, "NULL buffer");
    if (len >= (1ull << 32)) return fail(ENG_ERR_INVALID, "input too long");
    DevBuf in_buf, dig_buf;
    ST(in_buf.alloc(count * (len ? len : 1)));
    ST(dig_buf.alloc(count * 4));
    u64 *d_in = in_buf.p, *d_dig = dig_buf.p;
    if (len) CU(cudaMemcpyAsync(d_in, in_host, count * len * 8, cudaMemcpyHostToDevice, g.stream));
    // a zero-layer "tree" whose cap is the leaf digests: one sponge per row
    MerkleParams mp;
    mp.data = d_in; mp.row_stride = len; mp.col_stride = 1; mp.width = (u32)len; mp.num_leaves = count;
    mp.num_layers = 0; mp.digests = nullptr; mp.cap = d_dig;
    mp.noop_max = or_noop ? 4 : 0;  // hash_no_pad always runs the sponge (an empty input squeezes the zero state)
    eng_status s = ENG_OK;
    merkle_leaves_kernel<<<(unsigned)((count + 127) / 128), 128, 0, g.stream>>>(mp);
    g.launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) s = fail(ENG_ERR_CUDA, "merkle_leaves_kernel: %s", cudaGetErrorString(e));
    if (s == ENG_OK) s = d2h(out_host, d_dig, count * 32);
    return s;
}

eng_status eng_two_to_one(const uint64_t *pairs_host, size_t count, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (count == 0) return ENG_OK;
    if (!pairs_host || !out_host) return fail(ENG_ERR_INVALID, "NULL buffer");
    DevBuf d_in, d_out;
    ST(d_in.alloc(count * 8));
    ST(d_out.alloc(count * 4));
    CU(cudaMemcpyAsync(d_in.p, pairs_host, count * 64, cudaMemcpyHostToDevice, g.stream));
    two_to_one_kernel<<<(unsigned)((count + 127) / 128), 128, 0, g.stream>>>(d_in.p, d_out.p, count);
    g.launches++;
    CU(cudaGetLastError());
    return d2h(out_host, d_out.p, count * 32);
}

eng_status eng_batch_from_values(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                 int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!cols_host) return fail(ENG_ERR_INVALID, "cols_host is NULL");
    return make_batch(cols_host, nullptr, true, num_polys, log_n, rate_bits, blinding, blinding_seed, cap_height, out);
}
eng_status eng_batch_from_coeffs(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                 int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!cols_host) return fail(ENG_ERR_INVALID, "cols_host is NULL");
    return make_batch(cols_host, nullptr, false, num_polys, log_n, rate_bits, blinding, blinding_seed, cap_height, out);
}
eng_status eng_batch_from_values_dev(const uint64_t *values_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                     int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!values_dev) return fail(ENG_ERR_INVALID, "values_dev is NULL");
    return make_batch(nullptr, values_dev, true, num_polys, log_n, rate_bits, blinding, blinding_seed, cap_height, out);
}
eng_status eng_batch_from_coeffs_dev(const uint64_t *coeffs_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                                     int32_t blinding, uint64_t blinding_seed, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!coeffs_dev) return fail(ENG_ERR_INVALID, "coeffs_dev is NULL");
    return make_batch(nullptr, coeffs_dev, false, num_polys, log_n, rate_bits, blinding, blinding_seed, cap_height, out);
}

eng_status eng_lde_dev(const uint64_t *src_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits, int32_t is_values,
                       uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *lde_out_dev) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!src_dev || !coeffs_out_dev || !lde_out_dev) return fail(ENG_ERR_INVALID, "NULL buffer");
    if (num_polys == 0) return fail(ENG_ERR_INVALID, "no polynomials");
    if (log_n > 2 * NTT_MAX_LOGP || log_n + rate_bits > 32) return fail(ENG_ERR_INVALID, "size unsupported");
    const u64 n = (u64)1 << log_n;
    std::vector<NttLaunch> plan;
    if (is_values) {
        // the LDE output buffer is the four-step scratch of the iNTT
        if (!ntt_plan_intt(g.tables, src_dev, n, lde_out_dev, n, coeffs_out_dev, n, num_polys, log_n, plan))
            return fail(ENG_ERR_INVALID, "iNTT size unsupported");
        ST(launch_plan(plan));
    } else if (src_dev != coeffs_out_dev) {
        CU(cudaMemcpyAsync(coeffs_out_dev, src_dev, (size_t)num_polys * n * sizeof(u64), cudaMemcpyDeviceToDevice, g.stream));
    }
    plan.clear();
    if (!ntt_plan_lde(g.tables, coeffs_out_dev, n, lde_out_dev, num_polys, log_n, rate_bits, log_row_shards, plan))
        return fail(ENG_ERR_INVALID, "LDE shape unsupported (log_n=%u rate_bits=%u log_row_shards=%u)", log_n, rate_bits, log_row_shards);
    ST(launch_plan(plan));
    return ENG_OK;
}

// ---- fused exchange: the LDE's last pass stores straight into the row-shard owners' leaf matrices (NVLink P2P) ----
eng_status eng_lde_peer_dev(const uint64_t *src_dev, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits, int32_t is_values,
                            uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *scratch_dev, uint64_t *const *shard_out,
                            uint32_t first_shard) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!src_dev || !coeffs_out_dev || !scratch_dev || !shard_out) return fail(ENG_ERR_INVALID, "NULL buffer");
    if (num_polys == 0) return fail(ENG_ERR_INVALID, "no polynomials");
    if (log_n > 2 * NTT_MAX_LOGP || log_n + rate_bits > 32) return fail(ENG_ERR_INVALID, "size unsupported");
    if (((u64)1 << log_row_shards) > NTT_MAX_SHARDS) return fail(ENG_ERR_INVALID, "more than %d row shards", NTT_MAX_SHARDS);
    for (u32 gi = 0; gi < (1u << log_row_shards); gi++)
        if (!shard_out[gi]) return fail(ENG_ERR_INVALID, "shard_out[%u] is NULL", gi);
    const u64 n = (u64)1 << log_n, L = n << rate_bits, rows_per_shard = L >> log_row_shards;
    const u32 G = 1u << log_row_shards;
    // Column chunks (option lde_peer_chunk_cols, 0 = the whole shard at once): iNTT and LDE alternate chunk by chunk as in
    // the host-column pipeline, which spreads the peer stores of the last pass over the whole transform instead of
    // issuing all of them at the end (every rank of the box stores at the same time).
    const u32 chunk = g.opt_peer_chunk_cols > 0 && (u32)g.opt_peer_chunk_cols < num_polys ? (u32)g.opt_peer_chunk_cols : num_polys;
    if (!is_values && src_dev != coeffs_out_dev)
        CU(cudaMemcpyAsync(coeffs_out_dev, src_dev, (size_t)num_polys * n * sizeof(u64), cudaMemcpyDeviceToDevice, g.stream));
    std::vector<NttLaunch> plan;
    for (u32 c0 = 0; c0 < num_polys; c0 += chunk) {
        const u32 cc = num_polys - c0 < chunk ? num_polys - c0 : chunk;
        u64 *co = coeffs_out_dev + (size_t)c0 * n, *sc = scratch_dev + (size_t)c0 * L;
        if (is_values) {
            plan.clear();
            if (!ntt_plan_intt(g.tables, src_dev + (size_t)c0 * n, n, sc, n, co, n, cc, log_n, plan)) return fail(ENG_ERR_INVALID, "iNTT size unsupported");
            ST(launch_plan(plan));
        }
        u64 *so[NTT_MAX_SHARDS];
        for (u32 gi = 0; gi < G; gi++) so[gi] = shard_out[gi] + (size_t)c0 * rows_per_shard;
        plan.clear();
        if (!ntt_plan_lde(g.tables, co, n, sc, cc, log_n, rate_bits, log_row_shards, plan, so, first_shard))
            return fail(ENG_ERR_INVALID, "LDE shape unsupported (log_n=%u rate_bits=%u log_row_shards=%u)", log_n, rate_bits, log_row_shards);
        ST(launch_plan(plan));
    }
    return ENG_OK;
}

// Same with HOST columns: the copy / iNTT / LDE pipeline of make_batch, chunk by chunk, with the fused peer stores.
eng_status eng_lde_peer_host(const uint64_t *const *cols_host, uint32_t num_polys, uint32_t log_n, uint32_t rate_bits,
                             int32_t is_values, uint32_t log_row_shards, uint64_t *coeffs_out_dev, uint64_t *scratch_dev,
                             uint64_t *const *shard_out, uint32_t first_shard) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!cols_host || !coeffs_out_dev || !scratch_dev || !shard_out) return fail(ENG_ERR_INVALID, "NULL buffer");
    if (num_polys == 0) return fail(ENG_ERR_INVALID, "no polynomials");
    if (log_n > 2 * NTT_MAX_LOGP || log_n + rate_bits > 32) return fail(ENG_ERR_INVALID, "size unsupported");
    const u32 G = 1u << log_row_shards;
    if (G > NTT_MAX_SHARDS) return fail(ENG_ERR_INVALID, "more than %d row shards", NTT_MAX_SHARDS);
    for (u32 gi = 0; gi < G; gi++)
        if (!shard_out[gi]) return fail(ENG_ERR_INVALID, "shard_out[%u] is NULL", gi);
    for (u32 c = 0; c < num_polys; c++)
        if (!cols_host[c]) return fail(ENG_ERR_INVALID, "column %u is NULL", c);
    const u64 n = (u64)1 << log_n, L = n << rate_bits, rows_per_shard = L >> log_row_shards;
    u64 per = env_bytes("ENG_H2D_CHUNK_BYTES", (size_t)32 << 20) / (n * sizeof(u64));
    const u32 chunk = (u32)(per < 1 ? 1 : (per > num_polys ? num_polys : per));
    CU(cudaEventRecord(g.fork_ev, g.stream));
    CU(cudaStreamWaitEvent(g.copy_stream, g.fork_ev, 0));
    // on failure the caller will release coeffs/scratch: no copy may still be in flight into them
    auto run = [&]() -> eng_status {
        std::vector<NttLaunch> plan;
        for (u32 c0 = 0; c0 < num_polys; c0 += chunk) {
            const u32 cc = num_polys - c0 < chunk ? num_polys - c0 : chunk;
            u64 *co = coeffs_out_dev + (size_t)c0 * n, *sc = scratch_dev + (size_t)c0 * L;
            ST(h2d_columns(cols_host + c0, cc, n, co));
            CU(cudaEventRecord(g.chunk_ev, g.copy_stream));
            CU(cudaStreamWaitEvent(g.stream, g.chunk_ev, 0));
            if (is_values) {
                plan.clear();
                if (!ntt_plan_intt(g.tables, co, n, sc, n, co, n, cc, log_n, plan)) return fail(ENG_ERR_INVALID, "iNTT size unsupported");
                ST(launch_plan(plan));
            }
            u64 *so[NTT_MAX_SHARDS];
            for (u32 gi = 0; gi < G; gi++) so[gi] = shard_out[gi] + (size_t)c0 * rows_per_shard;
            plan.clear();
            if (!ntt_plan_lde(g.tables, co, n, sc, cc, log_n, rate_bits, log_row_shards, plan, so, first_shard))
                return fail(ENG_ERR_INVALID, "LDE shape unsupported (log_n=%u rate_bits=%u log_row_shards=%u)", log_n, rate_bits, log_row_shards);
            ST(launch_plan(plan));
        }
        return ENG_OK;
    };
    const eng_status st = run();
    if (st != ENG_OK) cudaStreamSynchronize(g.copy_stream);
    return st;
}

// Exchange buffers live outside the stream-ordered pool (cudaMalloc) so that they can be exported over CUDA IPC.
eng_status eng_peer_buffer_alloc(uint64_t num_elems, uint64_t **dev_out, uint8_t handle_out[64]) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!dev_out || !handle_out || num_elems == 0) return fail(ENG_ERR_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, num_elems * sizeof(u64));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ENG_ERR_OOM, "cudaMalloc of %llu bytes failed: %s", (unsigned long long)(num_elems * 8), cudaGetErrorString(e)); }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return fail(ENG_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *dev_out = (uint64_t *)p;
    return ENG_OK;
}
eng_status eng_peer_buffer_open(const uint8_t handle[64], uint64_t **dev_out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!handle || !dev_out) return fail(ENG_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_out = (uint64_t *)p;
    return ENG_OK;
}
eng_status eng_peer_buffer_close(uint64_t *peer_ptr) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!peer_ptr) return ENG_OK;
    CU(cudaIpcCloseMemHandle(peer_ptr));
    return ENG_OK;
}
eng_status eng_peer_buffer_free(uint64_t *dev_ptr) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!dev_ptr) return ENG_OK;
    CU(cudaDeviceSynchronize());
    CU(cudaFree(dev_ptr));
    return ENG_OK;
}

eng_status eng_batch_free(eng_batch *b) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!b) return ENG_OK;
    ST(check_ready());
    destroy_batch(b);
    return ENG_OK;
}

static eng_status merkle_common(const u64 *data_dev, bool owned, u64 row_stride, u64 col_stride, uint64_t num_leaves,
                                uint32_t leaf_len, uint32_t cap_height, eng_batch **out, u64 *owned_buf) {
    u32 log_l = 0;
    while (((u64)1 << log_l) < num_leaves) log_l++;
    eng_batch *b = new eng_batch();
    b->leaf_len = leaf_len; b->cap_height = cap_height; b->num_leaves = num_leaves;
    b->degree_log = log_l; b->num_layers = log_l - cap_height;
    b->num_digests = 2 * (num_leaves - ((u64)1 << cap_height));
    if (owned) b->lde = owned_buf;
    b->leaf_data = data_dev; b->row_stride = row_stride; b->col_stride = col_stride;
    eng_status st;
    if ((st = dev_alloc(&b->digests, (size_t)b->num_digests * 4)) != ENG_OK || (st = dev_alloc(&b->cap, (size_t)4 << cap_height)) != ENG_OK ||
        (st = build_tree(b)) != ENG_OK || (st = collect_times(b, false, false)) != ENG_OK) {
        destroy_batch(b);
        return st;
    }
    *out = b;
    return ENG_OK;
}

static eng_status merkle_check(uint64_t num_leaves, uint32_t leaf_len, uint32_t cap_height, eng_batch **out) {
    ST(check_ready());
    if (!out) return fail(ENG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (num_leaves == 0 || (num_leaves & (num_leaves - 1))) return fail(ENG_ERR_INVALID, "number of leaves %llu is not a power of two (log2_strict)", (unsigned long long)num_leaves);
    u32 log_l = 0;
    while (((u64)1 << log_l) < num_leaves) log_l++;
    if (cap_height > log_l) return fail(ENG_ERR_INVALID, "cap_height %u > log2(leaves) %u (MerkleTree::new assertion)", cap_height, log_l);
    (void)leaf_len;
    return ENG_OK;
}

eng_status eng_merkle_new(const uint64_t *leaves_host, uint64_t num_leaves, uint32_t leaf_len, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(merkle_check(num_leaves, leaf_len, cap_height, out));
    if (!leaves_host && leaf_len) return fail(ENG_ERR_INVALID, "leaves_host is NULL");
    DevBuf buf;
    ST(buf.alloc((size_t)num_leaves * (leaf_len ? leaf_len : 1)));
    CU(cudaEventRecord(g.ev[0], g.stream));
    if (leaf_len) CU(cudaMemcpyAsync(buf.p, leaves_host, (size_t)num_leaves * leaf_len * 8, cudaMemcpyHostToDevice, g.stream));
    const u64 *data = buf.p;
    return merkle_common(data, true, leaf_len, 1, num_leaves, leaf_len, cap_height, out, buf.take());
}

eng_status eng_merkle_new_dev(const uint64_t *data_dev, uint64_t row_stride, uint64_t col_stride, uint64_t num_leaves,
                              uint32_t leaf_len, uint32_t cap_height, eng_batch **out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(merkle_check(num_leaves, leaf_len, cap_height, out));
    if (!data_dev && leaf_len) return fail(ENG_ERR_INVALID, "data_dev is NULL");
    return merkle_common(data_dev, false, row_stride, col_stride, num_leaves, leaf_len, cap_height, out, nullptr);
}

eng_status eng_batch_info(const eng_batch *b, eng_batch_info_t *out) {
    if (!b || !out) return fail(ENG_ERR_INVALID, "NULL argument");
    out->num_polys = b->num_polys; out->degree_log = b->degree_log; out->rate_bits = b->rate_bits;
    out->cap_height = b->cap_height; out->blinding = b->blinding; out->leaf_len = b->leaf_len;
    out->num_leaves = b->num_leaves; out->num_digests = b->num_digests;
    return ENG_OK;
}

eng_status eng_batch_cap(const eng_batch *b, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || !out_host) return fail(ENG_ERR_INVALID, "NULL argument");
    return d2h(out_host, b->cap, (size_t)32 << b->cap_height);
}

eng_status eng_batch_digests(const eng_batch *b, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || (!out_host && b->num_digests)) return fail(ENG_ERR_INVALID, "NULL argument");
    if (b->num_digests == 0) return ENG_OK;
    return d2h(out_host, b->digests, (size_t)b->num_digests * 32);
}

eng_status eng_batch_coeffs(const eng_batch *b, uint32_t poly, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || !out_host) return fail(ENG_ERR_INVALID, "NULL argument");
    if (!b->coeffs || poly >= b->num_polys) return fail(ENG_ERR_INVALID, "polynomial index %u out of range (%u)", poly, b->num_polys);
    size_t n = (size_t)1 << b->degree_log;
    return d2h(out_host, b->coeffs + (size_t)poly * n, n * 8);
}

static eng_status gather_rows(const eng_batch *b, uint64_t first, uint64_t count, uint32_t width, uint64_t *out_host) {
    if (count == 0 || width == 0) return ENG_OK;
    DevBuf tmp;
    ST(tmp.alloc((size_t)count * width));
    u64 total = count * width;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, g.stream>>>(b->leaf_data, b->row_stride, b->col_stride, width, first, count, tmp.p);
    g.launches++;
    CU(cudaGetLastError());
    return d2h(out_host, tmp.p, (size_t)total * 8);
}

eng_status eng_batch_leaves(const eng_batch *b, uint64_t first, uint64_t count, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || (!out_host && count)) return fail(ENG_ERR_INVALID, "NULL argument");
    if (first > b->num_leaves || count > b->num_leaves - first) return fail(ENG_ERR_INVALID, "leaf range [%llu, +%llu) out of bounds", (unsigned long long)first, (unsigned long long)count);
    return gather_rows(b, first, count, b->leaf_len, out_host);
}

eng_status eng_batch_lde_values(const eng_batch *b, uint64_t index, uint64_t step, uint64_t *out_host) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || !out_host) return fail(ENG_ERR_INVALID, "NULL argument");
    u32 bits = b->degree_log + b->rate_bits;
    u64 i = index * step;
    if (i >= b->num_leaves) return fail(ENG_ERR_INVALID, "index*step = %llu out of bounds", (unsigned long long)i);
    u64 r = 0;
    for (u32 k = 0; k < bits; k++) r |= ((i >> k) & 1) << (bits - 1 - k);
    return gather_rows(b, r, 1, b->num_polys, out_host);
}

eng_status eng_batch_merkle_path(const eng_batch *b, uint64_t leaf_index, uint64_t *siblings_host, uint32_t *num_siblings) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    ST(check_ready());
    if (!b || !num_siblings) return fail(ENG_ERR_INVALID, "NULL argument");
    if (leaf_index >= b->num_leaves) return fail(ENG_ERR_INVALID, "leaf index %llu out of bounds", (unsigned long long)leaf_index);
    *num_siblings = b->num_layers;
    if (b->num_layers == 0) return ENG_OK;
    if (!siblings_host) return fail(ENG_ERR_INVALID, "siblings_host is NULL");
    // sibling of node g at layer i is node g^1 of the same layer
    u64 g_idx = leaf_index;
    for (u32 i = 0; i < b->num_layers; i++) {
        u64 pos = merkle_digest_pos(b->num_layers, i, g_idx ^ 1);
        CU(cudaMemcpyAsync(siblings_host + 4 * i, b->digests + 4 * pos, 32, cudaMemcpyDeviceToHost, g.stream));
        g_idx >>= 1;
    }
    CU(cudaStreamSynchronize(g.stream));
    return ENG_OK;
}

eng_status eng_batch_device_ptrs(const eng_batch *b, const uint64_t **lde_dev, const uint64_t **coeffs_dev,
                                 const uint64_t **digests_dev, const uint64_t **cap_dev) {
    if (!b) return fail(ENG_ERR_INVALID, "NULL argument");
    if (lde_dev) *lde_dev = b->leaf_data;
    if (coeffs_dev) *coeffs_dev = b->coeffs;
    if (digests_dev) *digests_dev = b->digests;
    if (cap_dev) *cap_dev = b->cap;
    return ENG_OK;
}

eng_status eng_batch_stage_ms(const eng_batch *b, float out[6]) {
    if (!b || !out) return fail(ENG_ERR_INVALID, "NULL argument");
    memcpy(out, b->stage_ms, sizeof(b->stage_ms));
    return ENG_OK;
}

}  // extern "C"

#include "engine_prover.inc"
#include "engine_plonk.inc"
#include "engine_verify.inc"
#include "engine_synth.inc"
