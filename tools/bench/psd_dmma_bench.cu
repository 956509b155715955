// Development tool (not shipped): the DMMA leaf kernel (csrc/poseidon_dmma.cuh) against merkle_leaves_kernel.
// Checks every digest word, then times both.  Build: see tools/bench/build_dmma.sh
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon_dmma.cuh"

#ifndef VARIANT_NAME
#define VARIANT_NAME "dmma"
#endif

int main(int argc, char **argv) {
    u32 log_rows = argc > 1 ? atoi(argv[1]) : 21;
    u32 width = argc > 2 ? atoi(argv[2]) : 135;
    u64 rows = 1ull << log_rows;
    u32 cap_h = 4;
    u64 *d, *dig0, *dig1, *cap;
    cudaMalloc(&d, rows * width * 8);
    size_t dig_words = 2 * (rows - (1u << cap_h)) * 4;
    cudaMalloc(&dig0, dig_words * 8);
    cudaMalloc(&dig1, dig_words * 8);
    cudaMalloc(&cap, 16 * 4 * 8);
    cudaMemset(dig0, 0, dig_words * 8);
    cudaMemset(dig1, 0, dig_words * 8);
    std::vector<u64> h(rows * width);
    u64 z = 88172645463325252ull;
    for (auto &v : h) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; v = z; }
    for (int i = 0; i < 64; i++) h[i * 7] = 0xFFFFFFFFFFFFFFFFull - i;   // non-canonical inputs
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    if (poseidon_upload_constants() != cudaSuccess || psd_dmma_upload_tables() != cudaSuccess) { printf("const upload failed\n"); return 1; }
    MerkleParams p{};
    p.data = d; p.row_stride = 1; p.col_stride = rows; p.width = width; p.noop_max = 4;
    p.num_leaves = rows; p.num_layers = log_rows - cap_h; p.cap = cap;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best[2] = {1e30f, 1e30f};
    for (int rep = 0; rep < 4; rep++) {
        for (int k = 0; k < 2; k++) {
            p.digests = k ? dig1 : dig0;
            cudaEventRecord(e0);
            if (k == 0) merkle_leaves_kernel<<<(unsigned)(rows / 128), 128>>>(p);
            else merkle_leaves_dmma_kernel<<<(unsigned)(rows / 128), 128>>>(p);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best[k]) best[k] = ms;
        }
    }
    cudaError_t e = cudaGetLastError();
    std::vector<u64> a(dig_words), b(dig_words);
    cudaMemcpy(a.data(), dig0, dig_words * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), dig1, dig_words * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < dig_words; i++) if (a[i] != b[i]) { if (!bad) first = i; bad++; }
    cudaFuncAttributes f0, f1;
    cudaFuncGetAttributes(&f0, merkle_leaves_kernel);
    cudaFuncGetAttributes(&f1, merkle_leaves_dmma_kernel);
    double perms = (double)rows * ((width + 7) / 8);
    printf("%-24s rows 2^%u width %u  scalar %.3f ms (%d regs)  dmma %.3f ms (%d regs, %zu B local)  x%.3f  %.1f Mperm/s  mismatching words %zu (first %zu) %s\n",
           VARIANT_NAME, log_rows, width, best[0], f0.numRegs, best[1], f1.numRegs, (size_t)f1.localSizeBytes, best[0] / best[1],
           perms / best[1] / 1e3, bad, first, cudaGetErrorString(e));
    if (bad) {
        // leaf digests sit at even positions of layer 0: print the first few leaves
        for (u64 j = 0; j < 4; j++) {
            u64 pos = 4 * merkle_digest_pos(p.num_layers, 0, j);
            printf(" leaf %llu: scalar %016llx %016llx | dmma %016llx %016llx\n", (unsigned long long)j, (unsigned long long)a[pos], (unsigned long long)a[pos + 1],
                   (unsigned long long)b[pos], (unsigned long long)b[pos + 1]);
        }
    }
    return bad ? 2 : 0;
}
