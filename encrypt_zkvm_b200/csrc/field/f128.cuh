// Device arithmetic in the reference's base field f128: M = 2^128 - 45*2^40 + 1
// (`winterfell::math::fields::f128::BaseElement`, prover/src/lib.rs:4,41).
// Elements are canonical (< M), 16 little-endian bytes in memory, moved with 128-bit loads/stores.
//
// No tensor cores: this is 32-bit integer-pipe work.  An element is four 32-bit limbs; the 256-bit product is
// accumulated in even/odd 64-bit columns so that every partial product is ONE IMAD.WIDE.U32 with carry-in/out
// (mad.lo.cc + madc.hi.cc pairs, which ptxas fuses), 16 of them per product.  The fold uses
//     2^128 = 45*2^40 - 1 = 11520*2^32 - 1 (mod M):
// hi*2^128 = (hi*11520) << 32 - hi, a limb-aligned shift, so the reduction is 4 more IMAD.WIDE and two
// carry chains, then a second tiny fold of the < 2^46 overflow.  ~58 SASS instructions per modmul
// (checked with cuobjdump), against ~120 for the 64-bit-limb formulation it replaces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ezk {
namespace dev {

struct fe {
    uint32_t a0, a1, a2, a3;
};

#define EZK_MOD_LO 0xFFFFD30000000001ULL
#define EZK_MOD_HI 0xFFFFFFFFFFFFFFFFULL
#define EZK_M1 0xFFFFD300u  // limb 1 of M (limb 0 = 1, limbs 2,3 = all ones)
#define EZK_K1 0x00002CFFu  // limb 1 of K = 2^128 - M = 45*2^40 - 1 (limb 0 = all ones, limbs 2,3 = 0)

__host__ __device__ __forceinline__ fe fe_make(uint64_t lo, uint64_t hi = 0) {
    fe r;
    r.a0 = (uint32_t)lo, r.a1 = (uint32_t)(lo >> 32), r.a2 = (uint32_t)hi, r.a3 = (uint32_t)(hi >> 32);
    return r;
}
__device__ __forceinline__ fe fe_zero() { return fe_make(0, 0); }
__device__ __forceinline__ fe fe_one() { return fe_make(1, 0); }
__device__ __forceinline__ bool fe_is_zero(fe a) { return (a.a0 | a.a1 | a.a2 | a.a3) == 0; }
__device__ __forceinline__ bool fe_eq(fe a, fe b) { return a.a0 == b.a0 && a.a1 == b.a1 && a.a2 == b.a2 && a.a3 == b.a3; }

__device__ __forceinline__ fe fe_from(uint4 v) {
    fe r;
    r.a0 = v.x, r.a1 = v.y, r.a2 = v.z, r.a3 = v.w;
    return r;
}
__device__ __forceinline__ uint4 fe_to(fe a) { return make_uint4(a.a0, a.a1, a.a2, a.a3); }
__device__ __forceinline__ fe fe_load(const uint4* p) { return fe_from(*p); }
__device__ __forceinline__ fe fe_ldg(const uint4* p) { return fe_from(__ldg(p)); }
__device__ __forceinline__ void fe_store(uint4* p, fe a) { *p = fe_to(a); }

// s += K (i.e. -M mod 2^128), dropping the carry
__device__ __forceinline__ void fe_add_k(fe& s) {
    asm("add.cc.u32 %0, %0, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 %1, %1, 0x2CFF;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(s.a0), "+r"(s.a1), "+r"(s.a2), "+r"(s.a3));
}

// Rarely taken tail of add / mul: `ov` = the value wrapped past 2^128 once; then canonicalise (s >= M -> s - M).
__device__ __forceinline__ fe fe_fix_rare(fe s, uint32_t ov) {
    if (ov) fe_add_k(s);
    if (s.a3 == 0xFFFFFFFFu && s.a2 == 0xFFFFFFFFu && (s.a1 > EZK_M1 || (s.a1 == EZK_M1 && s.a0 >= 1u))) fe_add_k(s);
    return s;
}

// a + b - [carry] * M: canonical unless the result lies in [M, 2^128), which needs limb 3 = all ones.
// Two encodings of the same operation (ptxas output: 12 and 10 instructions).  The masked one keeps everything in
// general registers; the predicated one turns the carry into the guard of four predicated additions.  Which one is
// faster depends on the kernel (predicate registers are scarce): see Arith below.
__device__ __forceinline__ fe fe_add_raw_masked(fe a, fe b) {
    fe s;
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=&r"(s.a0), "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3), "=&r"(c)
        : "r"(a.a0), "r"(a.a1), "r"(a.a2), "r"(a.a3), "r"(b.a0), "r"(b.a1), "r"(b.a2), "r"(b.a3));
    // carry out of 2^128: s - M = s + K - 2^128 (branch-free; about half of all additions take it)
    const uint32_t m0 = 0u - c, m1 = m0 & EZK_K1;
    asm("add.cc.u32 %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, %5;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(s.a0), "+r"(s.a1), "+r"(s.a2), "+r"(s.a3)
        : "r"(m0), "r"(m1));
    return s;
}
__device__ __forceinline__ fe fe_add_raw(fe a, fe b) {
    fe s;
    // carry out of 2^128: s - M = s + K - 2^128, applied by four additions predicated on the carry (about half of
    // all additions take it)
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 c;\n\t"
        "add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, %9;\n\t"
        "addc.cc.u32 %2, %6, %10;\n\t"
        "addc.cc.u32 %3, %7, %11;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "setp.ne.u32 p, c, 0;\n\t"
        "@p add.cc.u32 %0, %0, 0xFFFFFFFF;\n\t"
        "@p addc.cc.u32 %1, %1, 0x2CFF;\n\t"
        "@p addc.cc.u32 %2, %2, 0;\n\t"
        "@p addc.u32 %3, %3, 0;\n\t"
        "}"
        : "=&r"(s.a0), "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3)
        : "r"(a.a0), "r"(a.a1), "r"(a.a2), "r"(a.a3), "r"(b.a0), "r"(b.a1), "r"(b.a2), "r"(b.a3));
    return s;
}

// r = a + b (mod M), canonical inputs -> canonical output
__device__ __forceinline__ fe fe_add(fe a, fe b) {
    // the exact functions sit on rarely taken paths next to the hot loops; the register-only encoding keeps the
    // out-of-line redo functions from adding spills around their call sites
    fe s = fe_add_raw_masked(a, b);
    if (s.a3 == 0xFFFFFFFFu) s = fe_fix_rare(s, 0);  // only values >= 2^128 - 2^96 can still be >= M
    return s;
}

// Branch-free variants for straight-line butterfly code.  The rare tails (a sum in [M, 2^128); a product that
// wrapped past 2^128 in the last fold or landed above 2^128 - 2^96) are NOT fixed here: `rare` becomes
// 0xFFFFFFFF instead and the caller redoes the whole group with the exact functions.  Keeping the common path
// free of branches lets ptxas interleave the independent carry chains of neighbouring butterflies.
__device__ __forceinline__ fe fe_add_flag(fe a, fe b, uint32_t& rare) {
    fe s = fe_add_raw(a, b);
    rare = max(rare, s.a3);
    return s;
}
__device__ __forceinline__ fe fe_add_flag_masked(fe a, fe b, uint32_t& rare) {
    fe s = fe_add_raw_masked(a, b);
    rare = max(rare, s.a3);
    return s;
}

// r = a - b (mod M)
__device__ __forceinline__ fe fe_sub(fe a, fe b) {
    fe s;
    uint32_t m0;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=&r"(s.a0), "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3), "=&r"(m0)
        : "r"(a.a0), "r"(a.a1), "r"(a.a2), "r"(a.a3), "r"(b.a0), "r"(b.a1), "r"(b.a2), "r"(b.a3));
    // borrow: + M = - K (mod 2^128)
    const uint32_t m1 = m0 & EZK_K1;
    asm("sub.cc.u32 %0, %0, %4;\n\t"
        "subc.cc.u32 %1, %1, %5;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.u32 %3, %3, 0;"
        : "+r"(s.a0), "+r"(s.a1), "+r"(s.a2), "+r"(s.a3)
        : "r"(m0), "r"(m1));
    return s;
}

__device__ __forceinline__ fe fe_neg(fe a) { return fe_sub(fe_zero(), a); }

// s (128 bits) + top * 2^128 (top = p1:p0 < 2^46), modulo 2^128; ov = 1 when the sum wrapped (add K once more)
__device__ __forceinline__ fe fe_fold_top_raw(fe s, uint32_t p0, uint32_t p1, uint32_t& ov) {
    asm("{\n\t"
        ".reg .u32 v1, v2, q0, q1, q2;\n\t"
        "mul.lo.u32 v1, %5, 11520;\n\t"
        "mul.hi.u32 v2, %5, 11520;\n\t"
        "mad.lo.u32 v2, %6, 11520, v2;\n\t"  // top*11520 < 2^60
        "sub.cc.u32 q0, 0, %5;\n\t"          // q = (top*11520 << 32) - top  (>= 0)
        "subc.cc.u32 q1, v1, %6;\n\t"
        "subc.u32 q2, v2, 0;\n\t"
        "add.cc.u32 %0, %0, q0;\n\t"
        "addc.cc.u32 %1, %1, q1;\n\t"
        "addc.cc.u32 %2, %2, q2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "}"
        : "+r"(s.a0), "+r"(s.a1), "+r"(s.a2), "+r"(s.a3), "=r"(ov)
        : "r"(p0), "r"(p1));
    return s;
}

// The same fold for the flagged (branch-free) products: no carry capture.  top * (11520 * 2^32 - 1) < 2^92, so the
// sum can only wrap past 2^128 when limb 3 of `s` was all ones BEFORE the fold; the caller flags on that limb
// before and after (one three-input max) instead of materialising the carry.
__device__ __forceinline__ fe fe_fold_top_nowrap(fe s, uint32_t p0, uint32_t p1) {
    asm("{\n\t"
        ".reg .u32 v1, v2, q0, q1, q2;\n\t"
        "mul.lo.u32 v1, %4, 11520;\n\t"
        "mul.hi.u32 v2, %4, 11520;\n\t"
        "mad.lo.u32 v2, %5, 11520, v2;\n\t"
        "sub.cc.u32 q0, 0, %4;\n\t"
        "subc.cc.u32 q1, v1, %5;\n\t"
        "subc.u32 q2, v2, 0;\n\t"
        "add.cc.u32 %0, %0, q0;\n\t"
        "addc.cc.u32 %1, %1, q1;\n\t"
        "addc.cc.u32 %2, %2, q2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "+r"(s.a0), "+r"(s.a1), "+r"(s.a2), "+r"(s.a3)
        : "r"(p0), "r"(p1));
    return s;
}
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return max(a, max(b, c)); }

// ... -> canonical element
__device__ __forceinline__ fe fe_fold_top(fe s, uint32_t p0, uint32_t p1) {
    uint32_t ov;
    s = fe_fold_top_raw(s, p0, p1, ov);
    if (ov | (uint32_t)(s.a3 == 0xFFFFFFFFu)) s = fe_fix_rare(s, ov);
    return s;
}

// r[0..8) = a * b: even/odd column accumulation, 16 IMAD.WIDE.U32 with carry
__device__ __forceinline__ void fe_mul256(const fe& a, const fe& b, uint32_t (&r)[8]) {
    asm("{\n\t"
        ".reg .u32 e0, e1, e2, e3, e4, e5, e6, e7, o0, o1, o2, o3, o4, o5, o6;\n\t"
        // b0: even <- a0, a2 ; odd <- a1, a3
        "mul.lo.u32 e0, %8, %12;\n\t"
        "mul.hi.u32 e1, %8, %12;\n\t"
        "mul.lo.u32 e2, %10, %12;\n\t"
        "mul.hi.u32 e3, %10, %12;\n\t"
        "mul.lo.u32 o0, %9, %12;\n\t"
        "mul.hi.u32 o1, %9, %12;\n\t"
        "mul.lo.u32 o2, %11, %12;\n\t"
        "mul.hi.u32 o3, %11, %12;\n\t"
        // b1: odd += a0, a2 ; even += a1, a3
        "mad.lo.cc.u32 o0, %8, %13, o0;\n\t"
        "madc.hi.cc.u32 o1, %8, %13, o1;\n\t"
        "madc.lo.cc.u32 o2, %10, %13, o2;\n\t"
        "madc.hi.cc.u32 o3, %10, %13, o3;\n\t"
        "addc.u32 o4, 0, 0;\n\t"
        "mad.lo.cc.u32 e2, %9, %13, e2;\n\t"
        "madc.hi.cc.u32 e3, %9, %13, e3;\n\t"
        "madc.lo.cc.u32 e4, %11, %13, 0;\n\t"
        "madc.hi.u32 e5, %11, %13, 0;\n\t"
        // b2: even += a0, a2 ; odd += a1, a3
        "mad.lo.cc.u32 e2, %8, %14, e2;\n\t"
        "madc.hi.cc.u32 e3, %8, %14, e3;\n\t"
        "madc.lo.cc.u32 e4, %10, %14, e4;\n\t"
        "madc.hi.cc.u32 e5, %10, %14, e5;\n\t"
        "addc.u32 e6, 0, 0;\n\t"
        "mad.lo.cc.u32 o2, %9, %14, o2;\n\t"
        "madc.hi.cc.u32 o3, %9, %14, o3;\n\t"
        "madc.lo.cc.u32 o4, %11, %14, o4;\n\t"
        "madc.hi.u32 o5, %11, %14, 0;\n\t"
        // b3: odd += a0, a2 ; even += a1, a3
        "mad.lo.cc.u32 o2, %8, %15, o2;\n\t"
        "madc.hi.cc.u32 o3, %8, %15, o3;\n\t"
        "madc.lo.cc.u32 o4, %10, %15, o4;\n\t"
        "madc.hi.cc.u32 o5, %10, %15, o5;\n\t"
        "addc.u32 o6, 0, 0;\n\t"
        "mad.lo.cc.u32 e4, %9, %15, e4;\n\t"
        "madc.hi.cc.u32 e5, %9, %15, e5;\n\t"
        "madc.lo.cc.u32 e6, %11, %15, e6;\n\t"
        "madc.hi.u32 e7, %11, %15, 0;\n\t"
        // r = even + (odd << 32)
        "mov.u32 %0, e0;\n\t"
        "add.cc.u32 %1, e1, o0;\n\t"
        "addc.cc.u32 %2, e2, o1;\n\t"
        "addc.cc.u32 %3, e3, o2;\n\t"
        "addc.cc.u32 %4, e4, o3;\n\t"
        "addc.cc.u32 %5, e5, o4;\n\t"
        "addc.cc.u32 %6, e6, o5;\n\t"
        "addc.u32 %7, e7, o6;\n\t"
        "}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a.a0), "r"(a.a1), "r"(a.a2), "r"(a.a3), "r"(b.a0), "r"(b.a1), "r"(b.a2), "r"(b.a3));
}

// first fold of a 256-bit value: lo + hi*2^128 = lo + (hi*11520 << 32) - hi = s + (p1:p0) * 2^128, p1:p0 < 2^46
__device__ __forceinline__ void fe_reduce256_first(const uint32_t (&r)[8], fe& s, uint32_t& p0, uint32_t& p1) {
    asm("{\n\t"
        ".reg .u32 x1, x2, x3, x4, y2, y3, y4, y5;\n\t"
        // X = (h0*11520 + l2:l1) , (h2*11520 + l3) chained ; Y = h1*11520, h3*11520
        "mad.lo.cc.u32 x1, %10, 11520, %7;\n\t"
        "madc.hi.cc.u32 x2, %10, 11520, %8;\n\t"
        "madc.lo.cc.u32 x3, %12, 11520, %9;\n\t"
        "madc.hi.u32 x4, %12, 11520, 0;\n\t"
        "mul.lo.u32 y2, %11, 11520;\n\t"
        "mul.hi.u32 y3, %11, 11520;\n\t"
        "mul.lo.u32 y4, %13, 11520;\n\t"
        "mul.hi.u32 y5, %13, 11520;\n\t"
        "add.cc.u32 x2, x2, y2;\n\t"
        "addc.cc.u32 x3, x3, y3;\n\t"
        "addc.cc.u32 x4, x4, y4;\n\t"
        "addc.u32 y5, y5, 0;\n\t"
        // S = (l0, x1, x2, x3 | x4, y5) - h   (>= 0)
        "sub.cc.u32 %0, %6, %10;\n\t"
        "subc.cc.u32 %1, x1, %11;\n\t"
        "subc.cc.u32 %2, x2, %12;\n\t"
        "subc.cc.u32 %3, x3, %13;\n\t"
        "subc.cc.u32 %4, x4, 0;\n\t"
        "subc.u32 %5, y5, 0;\n\t"
        "}"
        : "=&r"(s.a0), "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3), "=&r"(p0), "=&r"(p1)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}

__device__ __forceinline__ fe fe_reduce256(const uint32_t (&r)[8]) {
    fe s;
    uint32_t p0, p1;
    fe_reduce256_first(r, s, p0, p1);
    return fe_fold_top(s, p0, p1);
}

__device__ __forceinline__ fe fe_mul(fe a, fe b) {
    uint32_t r[8];
    fe_mul256(a, b, r);
    return fe_reduce256(r);
}

// last fold of a flagged product.  MODE 1: flag on limb 3 before and after the fold with one three-input max (58
// instructions per product, one more register alive across the fold); MODE 2: the same test as two separate max
// operations (59, no extra register); MODE 0: capture the fold's carry instead (60).
template <int MODE>
__device__ __forceinline__ fe fe_fold_top_flag(fe s, uint32_t p0, uint32_t p1, uint32_t& rare) {
    if (MODE == 1) {
        const uint32_t before = s.a3;
        s = fe_fold_top_nowrap(s, p0, p1);
        rare = umax3(rare, before, s.a3);
    } else if (MODE == 2) {
        rare = max(rare, s.a3);
        s = fe_fold_top_nowrap(s, p0, p1);
        rare = max(rare, s.a3);
    } else {
        uint32_t ov;
        s = fe_fold_top_raw(s, p0, p1, ov);
        rare = max(rare, max(s.a3, 0u - ov));
    }
    return s;
}

// branch-free product, see fe_add_flag
template <int MODE = 1>
__device__ __forceinline__ fe fe_mul_flag(fe a, fe b, uint32_t& rare) {
    uint32_t r[8], p0, p1;
    fe s;
    fe_mul256(a, b, r);
    fe_reduce256_first(r, s, p0, p1);
    return fe_fold_top_flag<MODE>(s, p0, p1, rare);
}

__device__ __forceinline__ fe fe_sqr(fe a) { return fe_mul(a, a); }

// a * small (small < 2^32) = s + p0 * 2^128: 4 IMAD.WIDE
__device__ __forceinline__ void fe_mul_small_raw(fe a, uint32_t k, fe& s, uint32_t& p0) {
    asm("{\n\t"
        ".reg .u32 t1, t3;\n\t"
        "mul.lo.u32 %0, %5, %9;\n\t"
        "mul.hi.u32 %1, %5, %9;\n\t"
        "mul.lo.u32 %2, %7, %9;\n\t"
        "mul.hi.u32 %3, %7, %9;\n\t"
        "mul.lo.u32 t1, %6, %9;\n\t"
        "mul.hi.u32 %4, %8, %9;\n\t"
        "mul.lo.u32 t3, %8, %9;\n\t"
        "add.cc.u32 %1, %1, t1;\n\t"
        "madc.hi.cc.u32 %2, %6, %9, %2;\n\t"
        "addc.cc.u32 %3, %3, t3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "}"
        : "=&r"(s.a0), "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3), "=&r"(p0)
        : "r"(a.a0), "r"(a.a1), "r"(a.a2), "r"(a.a3), "r"(k));
}
__device__ __forceinline__ fe fe_mul_small(fe a, uint32_t k) {
    fe s;
    uint32_t p0;
    fe_mul_small_raw(a, k, s, p0);
    return fe_fold_top(s, p0, 0);
}

// ---- multiplication by a table constant in precomputed form ----
// For a constant w the table holds W_i = w * 2^(32 i) mod M, i = 0..3 (64 bytes).  Then x * w = sum_i x_i * W_i:
// four 32 x 128-bit products that all sit at limb 0, so the sum is < 2^162 and needs only the small second fold
// (no 256-bit combine, no first fold): 16 IMAD.WIDE + ~22 ALU instructions instead of 21 + ~34.  Every product of
// the NTT kernels has a table operand (butterfly twiddles, 8-point DFT constants, coset and inter-pass factors).
struct fe_pre {
    fe w[4];
};
__device__ __forceinline__ fe_pre fe_pre_ldg(const uint4* p) {
    fe_pre r;
    r.w[0] = fe_ldg(p), r.w[1] = fe_ldg(p + 1), r.w[2] = fe_ldg(p + 2), r.w[3] = fe_ldg(p + 3);
    return r;
}
__device__ __forceinline__ fe_pre fe_pre_load(const uint4* p) {
    fe_pre r;
    r.w[0] = fe_load(p), r.w[1] = fe_load(p + 1), r.w[2] = fe_load(p + 2), r.w[3] = fe_load(p + 3);
    return r;
}

// s + (p1:p0) * 2^128 = sum_i x_i * W_i
__device__ __forceinline__ void fe_mul_pre_raw(const fe& x, const fe_pre& W, fe& s, uint32_t& p0, uint32_t& p1) {
    uint32_t e0, e1, e2, e3, e4, o0, o1, o2, o3, o4;
    // i = 0: plain products; even columns from W_i limbs 0, 2, odd columns from limbs 1, 3
    asm("mul.lo.u32 %0, %8, %9;\n\t"
        "mul.hi.u32 %1, %8, %9;\n\t"
        "mul.lo.u32 %2, %8, %11;\n\t"
        "mul.hi.u32 %3, %8, %11;\n\t"
        "mul.lo.u32 %4, %8, %10;\n\t"
        "mul.hi.u32 %5, %8, %10;\n\t"
        "mul.lo.u32 %6, %8, %12;\n\t"
        "mul.hi.u32 %7, %8, %12;"
        : "=&r"(e0), "=&r"(e1), "=&r"(e2), "=&r"(e3), "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3)
        : "r"(x.a0), "r"(W.w[0].a0), "r"(W.w[0].a1), "r"(W.w[0].a2), "r"(W.w[0].a3));
    e4 = 0, o4 = 0;
#define EZK_PRE_ROW(XI, WI)                                                                                   \
    asm("mad.lo.cc.u32 %0, %10, %11, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %10, %11, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                                 \
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                                 \
        "addc.u32 %4, %4, 0;\n\t"                                                                              \
        "mad.lo.cc.u32 %5, %10, %12, %5;\n\t"                                                                  \
        "madc.hi.cc.u32 %6, %10, %12, %6;\n\t"                                                                 \
        "madc.lo.cc.u32 %7, %10, %14, %7;\n\t"                                                                 \
        "madc.hi.cc.u32 %8, %10, %14, %8;\n\t"                                                                 \
        "addc.u32 %9, %9, 0;"                                                                                  \
        : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4)   \
        : "r"(XI), "r"((WI).a0), "r"((WI).a1), "r"((WI).a2), "r"((WI).a3))
    EZK_PRE_ROW(x.a1, W.w[1]);
    EZK_PRE_ROW(x.a2, W.w[2]);
    EZK_PRE_ROW(x.a3, W.w[3]);
#undef EZK_PRE_ROW
    // r = even + (odd << 32): limbs 0..3 -> s, limbs 4, 5 -> p0, p1 (p1 < 4)
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, %13, 0;"
        : "=&r"(s.a1), "=&r"(s.a2), "=&r"(s.a3), "=&r"(p0), "=&r"(p1)
        : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4));
    s.a0 = e0;
}

__device__ __forceinline__ fe fe_mul_pre(const fe& x, const fe_pre& W) {
    fe s;
    uint32_t p0, p1;
    fe_mul_pre_raw(x, W, s, p0, p1);
    return fe_fold_top(s, p0, p1);
}
template <int MODE = 1>
__device__ __forceinline__ fe fe_mul_pre_flag(const fe& x, const fe_pre& W, uint32_t& rare) {
    fe s;
    uint32_t p0, p1;
    fe_mul_pre_raw(x, W, s, p0, p1);
    return fe_fold_top_flag<MODE>(s, p0, p1, rare);
}

// branch-free a * small, see fe_add_flag
__device__ __forceinline__ fe fe_mul_small_flag(fe a, uint32_t k, uint32_t& rare);

// Arithmetic policy of a straight-line block.  FAST: branch-free operations that only record their rare tails in
// `rare`; the caller checks tainted() once and recomputes the block with the exact policy (a few times per proof).
// LEAN selects the shorter encodings (predicated additions, products flagged without a carry capture: -2 instructions
// each).  Measured per kernel on the B200 at 2^20: constraint kernel 5.46 -> 4.49 ms, final NTT pass 7.08 -> 6.90 ms,
// but the strided NTT pass 10.05 -> 11.11 ms (it is the one kernel that loses: more values alive per thread), so that
// pass keeps the register-only encodings (LEAN = false).
template <bool FAST, bool LEAN = true, int LEAN_MUL = LEAN ? 1 : 0>
struct Arith {
    uint32_t rare = 0;
    __device__ __forceinline__ fe add(fe a, fe b) {
        return FAST ? (LEAN ? fe_add_flag(a, b, rare) : fe_add_flag_masked(a, b, rare)) : fe_add(a, b);
    }
    __device__ __forceinline__ fe sub(fe a, fe b) { return fe_sub(a, b); }
    __device__ __forceinline__ fe mul(fe a, fe b) { return FAST ? fe_mul_flag<LEAN_MUL>(a, b, rare) : fe_mul(a, b); }
    __device__ __forceinline__ fe mul_pre(const fe& a, const fe_pre& w) {
        return FAST ? fe_mul_pre_flag<LEAN_MUL>(a, w, rare) : fe_mul_pre(a, w);
    }
    __device__ __forceinline__ fe sqr(fe a) { return mul(a, a); }
    __device__ __forceinline__ fe cube(fe a) { return mul(mul(a, a), a); }
    __device__ __forceinline__ fe mul_small(fe a, uint32_t k) { return FAST ? fe_mul_small_flag(a, k, rare) : fe_mul_small(a, k); }
    __device__ __forceinline__ bool tainted() const { return FAST && rare == 0xFFFFFFFFu; }
    __device__ __forceinline__ void checkpoint() {}
};

// Same arithmetic, but checkpoint() is a CTA barrier: in a kernel whose straight-line code is larger than the
// instruction cache, keeping all warps of a CTA inside the same stretch of code turns per-warp instruction fetches
// into shared ones.  Every thread of the CTA must execute the same checkpoints.
struct ArithLockstep : Arith<true> {
    __device__ __forceinline__ void checkpoint() { __syncthreads(); }
};

__device__ __forceinline__ fe fe_mul_small_flag(fe a, uint32_t k, uint32_t& rare) {
    fe s;
    uint32_t p0;
    fe_mul_small_raw(a, k, s, p0);
    return fe_fold_top_flag<1>(s, p0, 0, rare);
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
