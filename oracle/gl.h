/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the product
 * path (eth-lc-plonky2_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED (except Poseidon): the arithmetic restated here lives in the third-party crates
 * plonky2 0.1.4 / plonky2_field 0.1.1 @ git 666f31517353b29b3d847c6e18b26c9be8bf060b
 * (pinned at /root/reference/Cargo.lock:2347-2350, 2424-2427), whose source is NOT under /root/reference and
 * cannot be fetched or compiled here (no Rust, no network).  This file restates the published algorithm;
 * the reference's own call sites are /root/reference/eth-lc-plonky2/src/main.rs:227,230.
 *
 * Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension F_p[X]/(X^2 - 7)
 * [DEP plonky2_field:goldilocks_field.rs, extension/quadratic.rs, goldilocks_extensions.rs].
 * All oracle values are kept CANONICAL (< p); inputs are canonicalised on entry.
 */
#ifndef ORACLE_GL_H
#define ORACLE_GL_H
#include <stdint.h>
#include <stddef.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL                     /* 2^64 mod p */
#define GL_GENERATOR 7ULL                        /* MULTIPLICATIVE_GROUP_GENERATOR = coset_shift() */
#define GL_TWO_ADICITY 32
#define GL_POWER_OF_TWO_GENERATOR 1753635133440165772ULL /* 7^((p-1)/2^32) */
#define GL_W 7ULL                                /* extension: X^2 = 7 */

static inline u64 gl_canon(u64 x) { return x >= GL_P ? x - GL_P : x; }

static inline u64 gl_add(u64 a, u64 b) { /* canonical in, canonical out; branch-free (the carry is a coin flip) */
    u64 s = a + b;
    u64 m = 0 - (u64)((s < a) | (s >= GL_P));
    return s - (GL_P & m);
}
static inline u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    return d + (GL_P & (0 - (u64)(a < b)));
}
static inline u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }

/* x = lo + hi*2^64, hi = hh*2^32 + hl:  2^64 = eps, 2^96 = -1  =>  x = lo - hh + hl*eps  (A.1) */
static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0;
    u64 borrow = __builtin_sub_overflow(lo, hh, &t0);
    t0 -= (0 - borrow) & GL_EPS;                /* borrow: +p  == -eps mod 2^64 */
    u64 t1 = hl * GL_EPS;                       /* < 2^64 */
    u64 t2;
    u64 carry = __builtin_add_overflow(t0, t1, &t2);
    t2 += (0 - carry) & GL_EPS;                 /* carry: -p == +eps mod 2^64 */
    return gl_canon(t2);
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static inline u64 gl_sqr(u64 a) { return gl_mul(a, a); }

static inline u64 gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_sqr(b); e >>= 1; }
    return r;
}
static inline u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

/* primitive_root_of_unity(k) = POWER_OF_TWO_GENERATOR^(2^(32-k)) [DEP plonky2_field:types.rs] */
static inline u64 gl_root_of_unity(int k) {
    u64 g = GL_POWER_OF_TWO_GENERATOR;
    for (int i = k; i < GL_TWO_ADICITY; i++) g = gl_sqr(g);
    return g;
}

static inline size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* ---- quadratic extension, flatten order [a0, a1] ---- */
typedef struct { u64 a, b; } gl2;
static inline gl2 gl2_make(u64 a, u64 b) { gl2 r = {a, b}; return r; }
static inline gl2 gl2_from(u64 a) { gl2 r = {a, 0}; return r; }
static inline gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
static inline gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
static inline gl2 gl2_neg(gl2 x) { return gl2_make(gl_neg(x.a), gl_neg(x.b)); }
static inline gl2 gl2_mul(gl2 x, gl2 y) {
    return gl2_make(gl_add(gl_mul(x.a, y.a), gl_mul(GL_W, gl_mul(x.b, y.b))),
                    gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
static inline gl2 gl2_scale(gl2 x, u64 s) { return gl2_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
static inline gl2 gl2_inv(gl2 x) { /* 1/(a+bX) = (a-bX)/(a^2-7b^2) */
    u64 n = gl_sub(gl_sqr(x.a), gl_mul(GL_W, gl_sqr(x.b)));
    u64 ni = gl_inv(n);
    return gl2_make(gl_mul(x.a, ni), gl_mul(gl_neg(x.b), ni));
}
static inline gl2 gl2_pow(gl2 b, u64 e) {
    gl2 r = gl2_from(1);
    while (e) { if (e & 1) r = gl2_mul(r, b); b = gl2_mul(b, b); e >>= 1; }
    return r;
}
static inline int gl2_eq(gl2 x, gl2 y) { return x.a == y.a && x.b == y.b; }

#endif
