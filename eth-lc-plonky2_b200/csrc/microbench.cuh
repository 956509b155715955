// Integer-pipe issue-rate microbenchmarks: the measured denominators for the Poseidon roofline
// (MEASURED_PEAKS.json has HBM and bf16 only).  Each thread runs 8 independent dependency chains so that the
// pipe, not latency, is the limit; the grid fills every SM with 2048 threads.
#pragma once
#include <stdint.h>

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8];
    uint64_t w[8];
    uint32_t m = seed | 1u, c = threadIdx.x + 1;
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed + k * 77u + threadIdx.x; w[k] = a[k]; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(c));        // IMAD
                if (KIND == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a[k]), "r"(m));  // IMAD.WIDE
                if (KIND == 2) asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(a[k]) : "r"(m), "r"(c));  // 2 alu ops
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32);
    if (r == 0x12345678u) out[0] = r;  // keeps the chains live
}
