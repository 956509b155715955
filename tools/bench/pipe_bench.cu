// Issue-rate microbenchmarks for the integer pipes of sm_100a (development tool).  Each kind is a loop body of
// known SASS composition (check with cuobjdump); the program prints time per loop iteration per warp in cycles.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;

#define W(k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((u32)w[((k) + 1) & 7]), "r"(m))
#define L(k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(m), "r"(c))
#define S(k) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[k]) : "r"(b[k]))
#define C2(k) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[k]), "+r"(b[k]) : "r"(m), "r"(c))
#define P2(k) asm volatile("{.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\tselp.u32 %0, %2, %0, p;}" : "+r"(a[k]) : "r"(m), "r"(c))

template <int KIND>
__global__ void __launch_bounds__(256) kern(u32 *out, u32 seed, int iters) {
    u32 a[8], b[8]; u64 w[8];
    u32 m = seed | 1u, c = threadIdx.x + 1;
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed + k * 77u + threadIdx.x; b[k] = a[k] * 3u; w[k] = a[k]; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (KIND == 0) { W(0); W(1); W(2); W(3); W(4); W(5); W(6); W(7); W(0); W(1); W(2); W(3); W(4); W(5); W(6); W(7); }
        if (KIND == 1) { L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        if (KIND == 2) { W(0); L(0); W(1); L(1); W(2); L(2); W(3); L(3); W(4); L(4); W(5); L(5); W(6); L(6); W(7); L(7); }
        if (KIND == 3) { W(0); L(0); L(1); W(1); L(2); L(3); W(2); L(4); L(5); W(3); L(6); L(7); W(4); L(0); L(1); W(5); L(2); L(3); }
        if (KIND == 4) { W(0); W(1); L(0); W(2); W(3); L(1); W(4); W(5); L(2); W(6); W(7); L(3); W(0); W(1); L(4); W(2); W(3); L(5); }
        if (KIND == 5) { C2(0); C2(1); C2(2); C2(3); C2(4); C2(5); C2(6); C2(7); }
        if (KIND == 6) { S(0); S(1); S(2); S(3); S(4); S(5); S(6); S(7); S(0); S(1); S(2); S(3); S(4); S(5); S(6); S(7); }
        if (KIND == 7) { W(0); W(1); W(2); W(3); W(4); W(5); W(6); W(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        if (KIND == 8) { W(0); W(1); W(2); W(3); W(4); W(5); W(6); W(7); W(0); W(1); W(2); W(3); W(4); W(5); W(6); W(7);
                         L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        if (KIND == 9) { P2(0); P2(1); P2(2); P2(3); P2(4); P2(5); P2(6); P2(7); }
        if (KIND == 10) { W(0); C2(0); W(1); C2(1); W(2); C2(2); W(3); C2(3); W(4); C2(4); W(5); C2(5); W(6); C2(6); W(7); C2(7); }
    }
    u32 r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k] ^ b[k] ^ (u32)w[k] ^ (u32)(w[k] >> 32);
    if (r == 0x12345678u) out[0] = r;
}

template <int KIND> void run(const char *name, int body_instrs, int warps_per_smsp) {
    u32 *d; cudaMalloc(&d, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int iters = 20000, threads = 256, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); kern<KIND><<<blocks, threads>>>(d, 1234u + rep, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = best * 1e-3 * clk * 1e3;   // at max clock
    double per_iter_per_smsp = cycles / iters;  // cycles per loop iteration for warps_per_smsp warps
    printf("%-28s warps/SMSP %2d  body %2d instr  %.2f cycles/iter/warp  IPC/SMSP %.3f\n", name, warps_per_smsp, body_instrs,
           per_iter_per_smsp / warps_per_smsp, body_instrs * warps_per_smsp / per_iter_per_smsp);
    cudaFree(d);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("16 IMAD.WIDE", 16, w);
        run<1>("16 LOP3", 16, w);
        run<2>("8x(WIDE,LOP3)", 16, w);
        run<3>("6x(WIDE,LOP3,LOP3)", 18, w);
        run<4>("6x(WIDE,WIDE,LOP3)", 18, w);
        run<5>("8x(IADD3.cc,IADD3.X)", 16, w);
        run<6>("16 SHF", 16, w);
        run<7>("8 WIDE then 8 LOP3", 16, w);
        run<8>("16 WIDE then 16 LOP3", 32, w);
        run<9>("8x(ISETP,SEL)", 16, w);
        run<10>("8x(WIDE,IADD3.cc,IADD3.X)", 24, w);
    }
    return 0;
}
