// Issue-rate microbenchmarks, part 2: the FP64 pipe of sm_100a and how it overlaps with the integer pipes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
__constant__ double c_k[16];

#define D(k) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[k]) : "d"(g[k]), "d"(h))
#define DC(k) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[k]) : "d"(g[k]), "d"(c_k[k]))
#define A(k) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[k]) : "d"(h))
#define W(k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((u32)w[((k) + 1) & 7]), "r"(m))
#define L(k) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(m), "r"(c))
#define I(k) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(m), "r"(c))
#define H(k) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b[k]) : "r"(m))
#define CV(k) asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(f[k]) : "r"(a[k]))

template <int KIND>
__global__ void __launch_bounds__(128) kern(u32 *out, u32 seed, int iters) {
    u32 a[8], b[8]; u64 w[8]; double f[8], g[8];
    u32 m = seed | 1u, c = threadIdx.x + 1;
    double h = 1.0 + seed * 1e-9;
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed + k * 77u + threadIdx.x; b[k] = a[k] * 3u; w[k] = a[k]; f[k] = a[k]; g[k] = 1e-3 * k; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (KIND == 0) { D(0); D(1); D(2); D(3); D(4); D(5); D(6); D(7); D(0); D(1); D(2); D(3); D(4); D(5); D(6); D(7); }
        if (KIND == 1) { DC(0); DC(1); DC(2); DC(3); DC(4); DC(5); DC(6); DC(7); DC(0); DC(1); DC(2); DC(3); DC(4); DC(5); DC(6); DC(7); }
        if (KIND == 2) { D(0); L(0); D(1); L(1); D(2); L(2); D(3); L(3); D(4); L(4); D(5); L(5); D(6); L(6); D(7); L(7); }
        if (KIND == 3) { D(0); W(0); D(1); W(1); D(2); W(2); D(3); W(3); D(4); W(4); D(5); W(5); D(6); W(6); D(7); W(7); }
        if (KIND == 4) { D(0); I(0); D(1); I(1); D(2); I(2); D(3); I(3); D(4); I(4); D(5); I(5); D(6); I(6); D(7); I(7); }
        if (KIND == 5) { I(0); I(1); I(2); I(3); I(4); I(5); I(6); I(7); I(0); I(1); I(2); I(3); I(4); I(5); I(6); I(7); }
        if (KIND == 6) { H(0); H(1); H(2); H(3); H(4); H(5); H(6); H(7); H(0); H(1); H(2); H(3); H(4); H(5); H(6); H(7); }
        if (KIND == 7) { L(0); I(0); L(1); I(1); L(2); I(2); L(3); I(3); L(4); I(4); L(5); I(5); L(6); I(6); L(7); I(7); }
        if (KIND == 8) { A(0); A(1); A(2); A(3); A(4); A(5); A(6); A(7); A(0); A(1); A(2); A(3); A(4); A(5); A(6); A(7); }
        if (KIND == 9) { CV(0); CV(1); CV(2); CV(3); CV(4); CV(5); CV(6); CV(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        if (KIND == 10) { D(0); D(1); L(0); W(0); D(2); D(3); L(1); W(1); D(4); D(5); L(2); W(2); D(6); D(7); L(3); W(3); }
        if (KIND == 11) { D(0); L(0); L(1); D(1); L(2); L(3); D(2); L(4); L(5); D(3); L(6); L(7); D(4); L(0); L(1); D(5); L(2); L(3); }
        if (KIND == 12) { D(0); D(1); L(0); D(2); D(3); L(1); D(4); D(5); L(2); D(6); D(7); L(3); D(0); D(1); L(4); D(2); D(3); L(5); }
        if (KIND == 13) { D(0); D(1); D(2); D(3); D(4); D(5); D(6); D(7); L(0); L(1); L(2); L(3); L(4); L(5); L(6); L(7); }
        if (KIND == 14) { W(0); I(0); W(1); I(1); W(2); I(2); W(3); I(3); W(4); I(4); W(5); I(5); W(6); I(6); W(7); I(7); }
        if (KIND == 15) { D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); D(0); }
    }
    u32 r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k] ^ b[k] ^ (u32)w[k] ^ (u32)(w[k] >> 32) ^ (u32)__double2loint(f[k]);
    if (r == 0x12345678u) out[0] = r;
}

template <int KIND> void run(const char *name, int body_instrs, int warps_per_smsp) {
    u32 *d; cudaMalloc(&d, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int iters = 20000, threads = 128, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); kern<KIND><<<blocks, threads>>>(d, 1234u + rep, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = best * 1e-3 * clk * 1e3;
    double per_iter_per_smsp = cycles / iters;
    printf("%-28s warps/SMSP %2d  body %2d instr  %.2f cycles/iter/warp  IPC/SMSP %.3f\n", name, warps_per_smsp, body_instrs,
           per_iter_per_smsp / warps_per_smsp, body_instrs * warps_per_smsp / per_iter_per_smsp);
    cudaFree(d);
}

int main() {
    double k[16]; for (int i = 0; i < 16; i++) k[i] = 1.0 + i * 1e-6;
    cudaMemcpyToSymbol(c_k, k, sizeof(k));
    for (int w : {1, 2, 4, 8}) {
        run<0>("16 DFMA", 16, w);
        run<1>("16 DFMA const operand", 16, w);
        run<15>("16 DFMA dependent chain", 16, w);
        run<8>("16 DADD", 16, w);
        run<2>("8x(DFMA,LOP3)", 16, w);
        run<11>("6x(DFMA,LOP3,LOP3)", 18, w);
        run<12>("6x(DFMA,DFMA,LOP3)", 18, w);
        run<13>("8 DFMA then 8 LOP3", 16, w);
        run<3>("8x(DFMA,WIDE)", 16, w);
        run<4>("8x(DFMA,IMAD32)", 16, w);
        run<10>("4x(DFMA,DFMA,LOP3,WIDE)", 16, w);
        run<5>("16 IMAD32", 16, w);
        run<6>("16 IMAD.HI", 16, w);
        run<7>("8x(LOP3,IMAD32)", 16, w);
        run<14>("8x(WIDE,IMAD32)", 16, w);
        run<9>("8 I2F.F64.U32 + 8 LOP3", 16, w);
    }
    return 0;
}
