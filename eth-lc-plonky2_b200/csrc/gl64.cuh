// Goldilocks field p = 2^64 - 2^32 + 1 on the 32-bit integer pipes of sm_100a.
//
// Replaces plonky2_field::goldilocks_field::GoldilocksField (dep plonky2_field 0.1.1, pinned at
// /root/reference/Cargo.lock:2424-2427).  Same value semantics: an element is a u64 that is
// NOT necessarily canonical; every routine here accepts any u64 and returns some representative < 2^64 of the
// exact residue.  Canonical form is produced only at the boundary (gl_canon) because equality, hashing to
// bytes and serialisation use it.
//
// Cost model (B300_MICROARCH.md): IMAD(.WIDE) issues on the fma pipe, IADD3/ISETP/SEL/LOP3 on the alu pipe,
// each 16 lanes/clk/SMSP; one SMSP issues one warp-instruction per clock.  A 64x64 product is 4 IMAD.WIDE;
// the reduction uses 2^64 = 2^32 - 1 and 2^96 = -1 (mod p), i.e. no multiplications, only carry chains.
#pragma once
#include <stdint.h>

// The arithmetic and the kernel bodies are written as __host__ __device__ functions so that tests/emu can
// replay the exact index math of every kernel on the CPU of the (GPU-less) development container.  That replay
// is a test harness only: the shipped library is built by nvcc for sm_100a and has no CPU path.
#ifndef __CUDACC__
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#define GL_HD __host__ __device__ __forceinline__

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFu

GL_HD u64 gl_canon(u64 x) { return x >= GL_P ? x - GL_P : x; }

// a + b (mod p) for ANY a, b < 2^64.  Two wrap corrections are needed only when both are non-canonical.
GL_HD u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    u64 c = s < a ? (u64)GL_EPS : 0;
    u64 t = s + c;
    if (t < s) t += GL_EPS;
    return t;
}
// a + b where b is known canonical (or a + b < 2^65 - 2^32): one wrap correction.
GL_HD u64 gl_add_c(u64 a, u64 b) {
    u64 s = a + b;
    return s + (s < a ? (u64)GL_EPS : 0);
}
// a - b (mod p) for ANY a, b.
GL_HD u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    u64 c = a < b ? (u64)GL_EPS : 0;
    u64 t = d - c;
    if (t > d) t -= GL_EPS;
    return t;
}
// a - b where b is known canonical: one wrap correction.
GL_HD u64 gl_sub_c(u64 a, u64 b) {
    u64 d = a - b;
    return d - (a < b ? (u64)GL_EPS : 0);
}
GL_HD u64 gl_neg(u64 a) { return gl_sub_c(0, gl_canon(a)); }

// x = lo + hi*2^64 = x0 + x1*2^32 + x2*2^64 + x3*2^96  ==>  x = (x1:x0) + x2*(2^32 - 1) - x3  (mod p).
// Device path: one 96-bit two's-complement carry chain V = (x1:x0) - x3 - x2 + (x2 << 32), V in (-2^33, 2^65), whose
// top limb v2 in {-1, 0, 1} is folded back as v2 * (2^32 - 1); neither fold can wrap a second time.
GL_HD u64 gl_reduce128(u64 lo, u64 hi) {
    u32 x2 = (u32)hi, x3 = (u32)(hi >> 32);
#ifdef __CUDA_ARCH__
    u32 x0 = (u32)lo, x1 = (u32)(lo >> 32);
    u32 v0, v1;
    asm("{\n\t"
        ".reg .u32 v2, t, h;\n\t"
        "sub.cc.u32  %0, %2, %5;\n\t"
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32    v2, 0, 0;\n\t"
        "sub.cc.u32  %0, %0, %4;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32    v2, v2, 0;\n\t"
        "add.cc.u32  %1, %1, %4;\n\t"
        "addc.u32    v2, v2, 0;\n\t"
        "sub.u32     t, 0, v2;\n\t"     // low word of v2*eps  (1 -> 0xffffffff, -1 -> 1)
        "shr.s32     h, v2, 1;\n\t"     // high word of v2*eps (1 -> 0, -1 -> 0xffffffff)
        "add.cc.u32  %0, %0, t;\n\t"
        "addc.u32    %1, %1, h;\n\t"
        "}"
        : "=&r"(v0), "=&r"(v1)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    return ((u64)v1 << 32) | v0;
#else
    u64 r = lo - x3;
    if (lo < x3) r -= GL_EPS;
    u64 t = r + (u64)x2 * GL_EPS;          // x2*eps <= 2^64 - 2^33 + 1
    return t + (t < r ? (u64)GL_EPS : 0);  // cannot wrap twice
#endif
}

GL_HD u64 gl_mulhi64(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
GL_HD u64 gl_mul(u64 a, u64 b) { return gl_reduce128(a * b, gl_mulhi64(a, b)); }
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// a*b + c (mod p), c any u64
GL_HD u64 gl_mul_add(u64 a, u64 b, u64 c) {
    u64 lo = a * b, hi = gl_mulhi64(a, b);
    u64 s = lo + c;
    hi += (s < lo);  // hi <= 2^64 - 2^33 + 1 before, no overflow
    return gl_reduce128(s, hi);
}

GL_HD u64 gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
GL_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// x^7
GL_HD u64 gl_pow7(u64 x) {
    u64 x2 = gl_sqr(x);
    u64 x4 = gl_sqr(x2);
    u64 x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

// ---- quadratic extension F_p[X]/(X^2 - 7), flatten order [a0, a1] ----
struct gl2 {
    u64 a, b;
};
GL_HD gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
GL_HD gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_HD gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_HD gl2 gl2_mul(gl2 x, gl2 y) {
    u64 bb = gl_mul(x.b, y.b);
    u64 r0 = gl_mul_add(x.a, y.a, gl_mul(bb, 7));
    u64 r1 = gl_mul_add(x.a, y.b, gl_mul(x.b, y.a));
    return gl2_make(r0, r1);
}
GL_HD gl2 gl2_scale(gl2 x, u64 s) { return gl2_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
GL_HD gl2 gl2_canon(gl2 x) { return gl2_make(gl_canon(x.a), gl_canon(x.b)); }

// ---- host-side exact arithmetic (table construction only) ----
static inline u64 h_gl_mul(u64 a, u64 b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;
    u64 t1 = hl * (u64)GL_EPS;
    u64 t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
    return t2 >= GL_P ? t2 - GL_P : t2;
}
static inline u64 h_gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = h_gl_mul(r, b);
        b = h_gl_mul(b, b);
        e >>= 1;
    }
    return r;
}
static inline u64 h_gl_inv(u64 a) { return h_gl_pow(a, GL_P - 2); }
static inline u64 h_gl_root_of_unity(int k) {  // primitive_root_of_unity(k)
    u64 g = 1753635133440165772ULL;            // POWER_OF_TWO_GENERATOR, order 2^32
    for (int i = k; i < 32; i++) g = h_gl_mul(g, g);
    return g;
}
