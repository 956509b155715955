// Standalone micro-benchmark of the Poseidon permutation kernel variants (development tool; not shipped).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DVARIANT=... -o psd_bench psd_bench.cu
// Each thread runs a sponge over `chunks` absorbs of 8 elements from a column-major matrix, like the leaf hash.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../eth-lc-plonky2_b200/csrc/poseidon.cuh"

#ifndef BLOCK
#define BLOCK 128
#endif
#ifndef MINB
#define MINB 1
#endif

__global__ void __launch_bounds__(BLOCK, MINB) sponge_kernel(const u64 *data, u64 rows, u32 chunks, u64 *out) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
#pragma unroll 1
    for (u32 c = 0; c < chunks; c++) {
        const u64 *src = data + (u64)(8 * c) * rows + j;
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = src[(u64)i * rows];
        poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[4 * j + i] = gl_canon(s[i]);
}

int main(int argc, char **argv) {
    u64 rows = argc > 1 ? strtoull(argv[1], 0, 10) : (1ull << 21);
    u32 chunks = argc > 2 ? atoi(argv[2]) : 17;
    u64 *d, *o;
    cudaMalloc(&d, rows * chunks * 8 * 8);
    cudaMalloc(&o, rows * 32);
    std::vector<u64> h(rows * chunks * 8);
    u64 z = 88172645463325252ull;
    for (auto &v : h) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; v = z; }
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    if (poseidon_upload_constants() != cudaSuccess) { printf("const upload failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        sponge_kernel<<<(unsigned)((rows + BLOCK - 1) / BLOCK), BLOCK>>>(d, rows, chunks, o);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    std::vector<u64> ho(8);
    cudaMemcpy(ho.data(), o, 64, cudaMemcpyDeviceToHost);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, sponge_kernel);
    double perms = (double)rows * chunks;
    printf("%-40s block %d regs %d  %.3f ms  %.1f Mperm/s  frac_of_18.6T_imad %.3f  digest0 %016llx %s\n", VARIANT_NAME, BLOCK, fa.numRegs,
           best, perms / best / 1e3, perms * 6612 / (best * 1e-3) / 18.61e12, (unsigned long long)ho[0], cudaGetErrorString(e));
    return 0;
}
