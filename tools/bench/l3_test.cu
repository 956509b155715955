// Device-vs-host check of the lazy 96-bit helpers of ntt.cuh (development tool).
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../eth-lc-plonky2_b200/csrc/ntt.cuh"

template <int S> __host__ __device__ u64 one(u64 a, u64 b, u64 c, u64 d) {
    gl96 x = l3_sub(l3_add(l3_from(a), l3_from(b)), l3_add(l3_from(c), l3_from(d)));   // |x| < 2^66
    gl96 y = l3_shl<S>(x);
    gl96 z = l3_add(y, l3_sub(l3_from(d), l3_from(a)));
    return gl_canon(l3_reduce(z));
}
__host__ __device__ void all(u64 a, u64 b, u64 c, u64 d, u64 *o) {
    o[0] = one<12>(a, b, c, d); o[1] = one<24>(a, b, c, d); o[2] = one<36>(a, b, c, d); o[3] = one<48>(a, b, c, d);
    o[4] = one<60>(a, b, c, d); o[5] = one<72>(a, b, c, d); o[6] = one<84>(a, b, c, d); o[7] = one<4>(a, b, c, d);
    o[8] = one<8>(a, b, c, d); o[9] = one<16>(a, b, c, d); o[10] = one<20>(a, b, c, d); o[11] = one<28>(a, b, c, d);
    o[12] = gl_canon(l3_reduce(l3_sub(l3_from(a), l3_from(b))));
    o[13] = gl_canon(l3_reduce(l3_add(l3_from(a), l3_from(b))));
    gl96 x[16];
    for (int i = 0; i < 16; i++) x[i] = l3_from(a * (i + 1) + b * (i * i + 3) + (i & 1 ? c : d));
    ntt16_stage<1, false>(x); ntt16_stage<2, false>(x); ntt16_stage<3, false>(x); ntt16_stage<4, false>(x);
    for (int i = 0; i < 16; i++) o[14 + i] = gl_canon(l3_reduce(x[i]));
}
__global__ void k(const u64 *in, u64 *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) all(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3], out + 30 * i);
}
int main() {
    int n = 4096;
    std::vector<u64> in(4 * n), ho(30 * n), dout(30 * n);
    u64 z = 88172645463325252ull;
    for (auto &v : in) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; v = z; }
    in[0] = in[1] = in[2] = in[3] = ~0ull; in[4] = 0; in[5] = ~0ull; in[6] = 0; in[7] = ~0ull;
    for (int i = 0; i < n; i++) all(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3], &ho[30 * i]);
    u64 *di, *dd;
    cudaMalloc(&di, in.size() * 8); cudaMalloc(&dd, dout.size() * 8);
    cudaMemcpy(di, in.data(), in.size() * 8, cudaMemcpyHostToDevice);
    k<<<(n + 127) / 128, 128>>>(di, dd, n);
    cudaMemcpy(dout.data(), dd, dout.size() * 8, cudaMemcpyDeviceToHost);
    int bad[30] = {0};
    for (int i = 0; i < n; i++) for (int j = 0; j < 30; j++) if (ho[30 * i + j] != dout[30 * i + j]) bad[j]++;
    for (int j = 0; j < 30; j++) printf("%d:%d ", j, bad[j]);
    printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
