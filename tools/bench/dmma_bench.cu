// Does the FP64 tensor-core path (mma.sync.m8n8k4.f64, "DMMA") free the issue port for the Poseidon MDS layers?
// (VERDICT r1 task 5.)  Per warp one DMMA is 8 x 8 x 4 = 256 FMAs = 8 per lane: it would replace 8 DFMA issues.
// Measured here: issue cycles per DMMA per warp per SMSP (alone, 4 independent accumulator sets), the same interleaved
// with alu work (LOP3) and with DFMA, and the DFMA rate for reference.  FMA/clk/SM = 256 * warps-in-flight / cycles.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;

#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
#define DF(k) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[k]) : "d"(g[k]), "d"(h))
#define L(k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m), "r"(cc))
#define IW(k) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(y[k]) : "r"(m))

template <int KIND>
__global__ void __launch_bounds__(128) kern(double *out, u32 seed, int iters) {
    double c[8], f[8], g[8];
    u32 x[8];
    unsigned long long y[8];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * seed, h = 1.0 + seed * 1e-12;
    u32 m = seed | 1u, cc = threadIdx.x + 1;
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k] = k; f[k] = k + threadIdx.x; g[k] = 1e-3 * k; x[k] = seed + k; y[k] = seed * 3 + k + threadIdx.x; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (KIND == 0) {
            DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
            DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
        }
        if (KIND == 1) {   // 8 DMMA + 16 LOP3
            DMMA(c[0], c[1], a, b); { L(0); L(1); } DMMA(c[2], c[3], a, b); { L(2); L(3); }
            DMMA(c[4], c[5], a, b); { L(4); L(5); } DMMA(c[6], c[7], a, b); { L(6); L(7); }
            DMMA(c[0], c[1], a, b); { L(0); L(1); } DMMA(c[2], c[3], a, b); { L(2); L(3); }
            DMMA(c[4], c[5], a, b); { L(4); L(5); } DMMA(c[6], c[7], a, b); { L(6); L(7); }
        }
        if (KIND == 2) {   // 8 DMMA + 16 DFMA
            DMMA(c[0], c[1], a, b); DF(0); DF(1); DMMA(c[2], c[3], a, b); DF(2); DF(3);
            DMMA(c[4], c[5], a, b); DF(4); DF(5); DMMA(c[6], c[7], a, b); DF(6); DF(7);
            DMMA(c[0], c[1], a, b); DF(0); DF(1); DMMA(c[2], c[3], a, b); DF(2); DF(3);
            DMMA(c[4], c[5], a, b); DF(4); DF(5); DMMA(c[6], c[7], a, b); DF(6); DF(7);
        }
        if (KIND == 3) { DF(0); DF(1); DF(2); DF(3); DF(4); DF(5); DF(6); DF(7); DF(0); DF(1); DF(2); DF(3); DF(4); DF(5); DF(6); DF(7); }
        if (KIND == 5) {   // 8 DMMA + 16 IMAD.WIDE
            DMMA(c[0], c[1], a, b); { IW(0); IW(1); } DMMA(c[2], c[3], a, b); { IW(2); IW(3); }
            DMMA(c[4], c[5], a, b); { IW(4); IW(5); } DMMA(c[6], c[7], a, b); { IW(6); IW(7); }
            DMMA(c[0], c[1], a, b); { IW(0); IW(1); } DMMA(c[2], c[3], a, b); { IW(2); IW(3); }
            DMMA(c[4], c[5], a, b); { IW(4); IW(5); } DMMA(c[6], c[7], a, b); { IW(6); IW(7); }
        }
        if (KIND == 6) { IW(0); IW(1); IW(2); IW(3); IW(4); IW(5); IW(6); IW(7); IW(0); IW(1); IW(2); IW(3); IW(4); IW(5); IW(6); IW(7); }
        if (KIND == 7) {   // odd warps: 8 DMMA, even warps: 16 IMAD.WIDE (the two kinds of work come from different warps)
            if ((threadIdx.x >> 5) & 1) {
                DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
                DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
            } else { IW(0); IW(1); IW(2); IW(3); IW(4); IW(5); IW(6); IW(7); IW(0); IW(1); IW(2); IW(3); IW(4); IW(5); IW(6); IW(7); }
        }
        if (KIND == 8) {   // odd warps: 8 DMMA, even warps: 16 LOP3
            if ((threadIdx.x >> 5) & 1) {
                DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
                DMMA(c[0], c[1], a, b); DMMA(c[2], c[3], a, b); DMMA(c[4], c[5], a, b); DMMA(c[6], c[7], a, b);
            } else { L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        }
        if (KIND == 4) { L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
    }
    double r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r += c[k] + f[k] + (double)x[k] + (double)y[k];
    if (r == 0.123456789) out[0] = r;
}

template <int KIND> void run(const char *name, int n_dmma, int n_other, int warps_per_smsp) {
    double *d; cudaMalloc(&d, 64);
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000, threads = 128, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); kern<KIND><<<blocks, threads>>>(d, 77u + rep, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * clk * 1e3 / ((double)iters * warps_per_smsp);   // per loop body per warp per SMSP
    printf("%d warps/SMSP  %-34s %7.1f cycles / body", warps_per_smsp, name, cycles);
    if (n_dmma) printf("  = %5.2f cycles per DMMA slot (%d DMMA + %d other)  ->  %6.1f FP64 FMA/clk/SM from DMMA", cycles / n_dmma, n_dmma, n_other, 256.0 * n_dmma * 4 / cycles);
    else printf("  = %5.2f cycles per instruction", cycles / n_other);
    printf("\n");
    cudaFree(d);
}

int main() {
    for (int w = 1; w <= 8; w *= 2) {
        run<0>("8 DMMA", 8, 0, w);
        run<1>("8 DMMA + 16 LOP3", 8, 16, w);
        run<2>("8 DMMA + 16 DFMA", 8, 16, w);
        run<3>("16 DFMA", 0, 16, w);
        run<4>("16 LOP3", 0, 16, w);
        run<5>("8 DMMA + 16 IMAD.WIDE", 8, 16, w);
        run<6>("16 IMAD.WIDE", 0, 16, w);
        if (w >= 2) {
            run<7>("warps: half 8 DMMA, half 16 IMAD.W", 8, 16, w);
            run<8>("warps: half 8 DMMA, half 16 LOP3", 8, 16, w);
        }
    }
    printf("DFMA peak of the SM: 64 FMA/clk (2 cycles per warp instruction per SMSP).\n");
    return 0;
}
