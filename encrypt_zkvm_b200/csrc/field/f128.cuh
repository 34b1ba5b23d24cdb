// Device arithmetic in the reference's base field f128: M = 2^128 - 45*2^40 + 1
// (`winterfell::math::fields::f128::BaseElement`, prover/src/lib.rs:4,41).
// Elements are canonical (< M), 16 little-endian bytes in memory, moved with 128-bit loads/stores.
// 2^128 = 45*2^40 - 1 (mod M), so a 256-bit product folds with one multiply-by-45, shifts and adds;
// no tensor cores: this is 32-bit integer-pipe work (IMAD.WIDE + IADD3 carry chains).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ezk {
namespace dev {

struct fe {
    uint64_t lo, hi;
};

#define EZK_MOD_LO 0xFFFFD30000000001ULL
#define EZK_MOD_HI 0xFFFFFFFFFFFFFFFFULL

__host__ __device__ __forceinline__ fe fe_make(uint64_t lo, uint64_t hi = 0) {
    fe r;
    r.lo = lo, r.hi = hi;
    return r;
}
__device__ __forceinline__ fe fe_zero() { return fe_make(0, 0); }
__device__ __forceinline__ fe fe_one() { return fe_make(1, 0); }
__device__ __forceinline__ bool fe_is_zero(fe a) { return (a.lo | a.hi) == 0; }
__device__ __forceinline__ bool fe_eq(fe a, fe b) { return a.lo == b.lo && a.hi == b.hi; }

__device__ __forceinline__ fe fe_load(const uint4* p) {
    uint4 v = *p;
    return fe_make(((uint64_t)v.y << 32) | v.x, ((uint64_t)v.w << 32) | v.z);
}
__device__ __forceinline__ fe fe_ldg(const uint4* p) {
    uint4 v = __ldg(p);
    return fe_make(((uint64_t)v.y << 32) | v.x, ((uint64_t)v.w << 32) | v.z);
}
__device__ __forceinline__ void fe_store(uint4* p, fe a) {
    uint4 v;
    v.x = (uint32_t)a.lo, v.y = (uint32_t)(a.lo >> 32), v.z = (uint32_t)a.hi, v.w = (uint32_t)(a.hi >> 32);
    *p = v;
}

// r = a + b (mod M), canonical inputs -> canonical output
__device__ __forceinline__ fe fe_add(fe a, fe b) {
    uint64_t lo, hi, c;
    asm("add.cc.u64 %0, %3, %5;\n\t"
        "addc.cc.u64 %1, %4, %6;\n\t"
        "addc.u64 %2, 0, 0;"
        : "=&l"(lo), "=&l"(hi), "=&l"(c)
        : "l"(a.lo), "l"(a.hi), "l"(b.lo), "l"(b.hi));
    // subtract M when the sum overflowed 2^128 or is >= M  (M.hi is all ones)
    bool ge = c || (hi == EZK_MOD_HI && lo >= EZK_MOD_LO);
    if (ge) {
        // s - M = s + (2^128 - M) - 2^128 = s + (45*2^40 - 1), dropping the carry
        const uint64_t k = (45ULL << 40) - 1;
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, 0;"
            : "+l"(lo), "+l"(hi)
            : "l"(k));
    }
    return fe_make(lo, hi);
}

// r = a - b (mod M)
__device__ __forceinline__ fe fe_sub(fe a, fe b) {
    uint64_t lo, hi, bw;
    asm("sub.cc.u64 %0, %3, %5;\n\t"
        "subc.cc.u64 %1, %4, %6;\n\t"
        "subc.u64 %2, 0, 0;"
        : "=&l"(lo), "=&l"(hi), "=&l"(bw)
        : "l"(a.lo), "l"(a.hi), "l"(b.lo), "l"(b.hi));
    if (bw) {
        // + M = - (45*2^40 - 1) mod 2^128
        const uint64_t k = (45ULL << 40) - 1;
        asm("sub.cc.u64 %0, %0, %2;\n\t"
            "subc.u64 %1, %1, 0;"
            : "+l"(lo), "+l"(hi)
            : "l"(k));
    }
    return fe_make(lo, hi);
}

__device__ __forceinline__ fe fe_neg(fe a) { return fe_sub(fe_zero(), a); }

// reduce a 256-bit value (r3:r2:r1:r0, 64-bit words) modulo M
__device__ __forceinline__ fe fe_reduce256(uint64_t r0, uint64_t r1, uint64_t r2, uint64_t r3) {
    // hi * 45 -> t2:t1:t0 (134 bits)
    uint64_t t0 = r2 * 45ULL;
    uint64_t t1 = __umul64hi(r2, 45ULL);
    uint64_t t2;
    {
        uint64_t m_lo = r3 * 45ULL, m_hi = __umul64hi(r3, 45ULL);
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %3, 0;"
            : "+l"(t1), "=&l"(t2)
            : "l"(m_lo), "l"(m_hi));
    }
    // u = t << 40 (174 bits)
    uint64_t u0 = t0 << 40;
    uint64_t u1 = (t1 << 40) | (t0 >> 24);
    uint64_t u2 = (t2 << 40) | (t1 >> 24);
    // v = u - hi  (>= 0)
    asm("sub.cc.u64 %0, %0, %3;\n\t"
        "subc.cc.u64 %1, %1, %4;\n\t"
        "subc.u64 %2, %2, 0;"
        : "+l"(u0), "+l"(u1), "+l"(u2)
        : "l"(r2), "l"(r3));
    // s = lo + v  -> top:s1:s0, top < 2^47
    uint64_t s0, s1, top;
    asm("add.cc.u64 %0, %3, %5;\n\t"
        "addc.cc.u64 %1, %4, %6;\n\t"
        "addc.u64 %2, %7, 0;"
        : "=&l"(s0), "=&l"(s1), "=&l"(top)
        : "l"(r0), "l"(r1), "l"(u0), "l"(u1), "l"(u2));
    // second fold: top * (45*2^40 - 1) = (top*45 << 40) - top, < 2^93
    uint64_t w = top * 45ULL;  // < 2^53
    uint64_t w0 = w << 40, w1 = w >> 24;
    asm("sub.cc.u64 %0, %0, %2;\n\t"
        "subc.u64 %1, %1, 0;"
        : "+l"(w0), "+l"(w1)
        : "l"(top));
    uint64_t c;
    asm("add.cc.u64 %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, %4;\n\t"
        "addc.u64 %2, 0, 0;"
        : "+l"(s0), "+l"(s1), "=&l"(c)
        : "l"(w0), "l"(w1));
    const uint64_t k = (45ULL << 40) - 1;
    if (c) {  // wrapped past 2^128 once more: add 2^128 mod M (cannot wrap again: s < 2^93 here)
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, 0;"
            : "+l"(s0), "+l"(s1)
            : "l"(k));
    }
    if (s1 == EZK_MOD_HI && s0 >= EZK_MOD_LO) {  // final canonicalisation
        asm("add.cc.u64 %0, %0, %2;\n\t"
            "addc.u64 %1, %1, 0;"
            : "+l"(s0), "+l"(s1)
            : "l"(k));
    }
    return fe_make(s0, s1);
}

__device__ __forceinline__ fe fe_mul(fe a, fe b) {
    // 256-bit schoolbook product on 64-bit limbs (lowered by ptxas to IMAD.WIDE chains)
    uint64_t p0l = a.lo * b.lo, p0h = __umul64hi(a.lo, b.lo);
    uint64_t p1l = a.lo * b.hi, p1h = __umul64hi(a.lo, b.hi);
    uint64_t p2l = a.hi * b.lo, p2h = __umul64hi(a.hi, b.lo);
    uint64_t p3l = a.hi * b.hi, p3h = __umul64hi(a.hi, b.hi);
    uint64_t r0 = p0l, r1, r2, r3;
    asm("add.cc.u64 %0, %3, %4;\n\t"
        "addc.cc.u64 %1, %5, %6;\n\t"
        "addc.u64 %2, %7, 0;\n\t"
        "add.cc.u64 %0, %0, %8;\n\t"
        "addc.cc.u64 %1, %1, %9;\n\t"
        "addc.u64 %2, %2, 0;"
        : "=&l"(r1), "=&l"(r2), "=&l"(r3)
        : "l"(p0h), "l"(p1l), "l"(p1h), "l"(p3l), "l"(p3h), "l"(p2l), "l"(p2h));
    return fe_reduce256(r0, r1, r2, r3);
}

__device__ __forceinline__ fe fe_sqr(fe a) { return fe_mul(a, a); }

// a * small (small < 2^32), cheaper than a full product
__device__ __forceinline__ fe fe_mul_small(fe a, uint32_t s) {
    uint64_t r0 = a.lo * s, c0 = __umul64hi(a.lo, (uint64_t)s);
    uint64_t r1 = a.hi * s, r2 = __umul64hi(a.hi, (uint64_t)s);
    asm("add.cc.u64 %0, %0, %2;\n\t"
        "addc.u64 %1, %1, 0;"
        : "+l"(r1), "+l"(r2)
        : "l"(c0));
    return fe_reduce256(r0, r1, r2, 0);
}

__device__ __forceinline__ fe fe_pow(fe b, uint64_t e) {
    fe r = fe_one();
    while (e) {
        if (e & 1) r = fe_mul(r, b);
        b = fe_sqr(b);
        e >>= 1;
    }
    return r;
}

// a^(M-2); inv(0) = 0 like winter-math
static __device__ __noinline__ fe fe_inv(fe a) {
    // M - 2 = 2^128 - 45*2^40 - 1: hi word all ones, lo word 0xFFFFD2FFFFFFFFFF
    fe r = fe_one();
    const uint64_t e_lo = EZK_MOD_LO - 2, e_hi = EZK_MOD_HI;
    for (int i = 127; i >= 0; i--) {
        r = fe_sqr(r);
        uint64_t bit = i >= 64 ? (e_hi >> (i - 64)) & 1 : (e_lo >> i) & 1;
        if (bit) r = fe_mul(r, a);
    }
    return r;
}

// Two-level power tables: T[0][k] = b^k, T[1][k] = b^(k*2^14), k < 2^14  =>  b^e for e < 2^28
#define EZK_TAB_BITS 14
#define EZK_TAB_SIZE (1u << EZK_TAB_BITS)
#define EZK_ROOT_LOG 28  // tables of the primitive 2^28-th root of unity

__device__ __forceinline__ fe fe_tab_pow(const uint4* __restrict__ tab, uint32_t e) {
    uint32_t lo = e & (EZK_TAB_SIZE - 1), hi = e >> EZK_TAB_BITS;
    fe a = fe_ldg(tab + lo);
    if (hi == 0) return a;
    fe b = fe_ldg(tab + EZK_TAB_SIZE + hi);
    if (lo == 0) return b;
    return fe_mul(a, b);
}

// w_N^e where N = 2^log_n <= 2^28, from the 2^28-th root tables (e taken mod N)
__device__ __forceinline__ fe fe_root_pow(const uint4* __restrict__ tab, uint32_t log_n, uint64_t e) {
    uint32_t ee = (uint32_t)(e & ((1ull << log_n) - 1)) << (EZK_ROOT_LOG - log_n);
    return fe_tab_pow(tab, ee);
}

}  // namespace dev
}  // namespace ezk
