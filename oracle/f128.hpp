// ORACLE (test infrastructure, not product): the reference prover's base field, restated on the CPU.
//
// The reference fixes `BaseField = winterfell::math::fields::f128::BaseElement`
// (prover/src/lib.rs:4,41; air/src/lib.rs:6).  That type lives in the un-vendored crate
// winter-math 0.9.0 (Cargo.lock:573-580); its published definition is restated here:
//   M = 2^128 - 45*2^40 + 1, canonical (non-Montgomery) u128 values, 16 little-endian bytes,
//   GENERATOR = 3, TWO_ADICITY = 40, TWO_ADIC_ROOT_OF_UNITY = 23953097886125630542083529559205016746,
//   inv(0) = 0.
// The constants are cross-checked against the reference itself in tests/test_oracle_field.py
// (crypto/src/rescue.rs:194-233: 3*INV_ALPHA = 1 mod M-1 and MDS*INV_MDS = I only hold for this M).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;

static inline constexpr u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

// M = 2^128 - 45*2^40 + 1
static constexpr u128 MOD = mk128(0xFFFFFFFFFFFFFFFFULL, 0xFFFFD30000000001ULL);
// 2^128 mod M = 45*2^40 - 1
static constexpr u128 C128 = ((u128)45 << 40) - 1;
static constexpr u128 GENERATOR = 3;
static constexpr unsigned TWO_ADICITY = 40;
// 23953097886125630542083529559205016746
static constexpr u128 TWO_ADIC_ROOT = mk128(0x120532E7B364080AULL, 0x86B8723E1920F4AAULL);

static inline u128 fadd(u128 a, u128 b) {
    u128 s = a + b;
    if (s < a || s >= MOD) s -= MOD;  // a,b < M < 2^128: one conditional subtraction
    return s;
}
static inline u128 fsub(u128 a, u128 b) { return a >= b ? a - b : a + (MOD - b); }
static inline u128 fneg(u128 a) { return a == 0 ? 0 : MOD - a; }

// full 256-bit product, then two folds with 2^128 = C128 (mod M)
static inline u128 fmul(u128 a, u128 b) {
    uint64_t a0 = (uint64_t)a, a1 = (uint64_t)(a >> 64), b0 = (uint64_t)b, b1 = (uint64_t)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    u128 lo = ((u128)(uint64_t)mid << 64) | (uint64_t)p00;
    u128 hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
    // hi * C128 as a 192-bit number: t_hi:t_lo
    uint64_t h0 = (uint64_t)hi, h1 = (uint64_t)(hi >> 64);
    u128 q0 = (u128)h0 * (uint64_t)C128;  // < 2^110
    u128 q1 = (u128)h1 * (uint64_t)C128;  // < 2^110
    u128 t_lo = q0 + (q1 << 64);
    u128 t_hi = (q1 >> 64) + (t_lo < q0 ? 1 : 0);  // < 2^47
    unsigned carries = 0;
    u128 s = lo + t_lo;
    carries += (s < lo);
    u128 f2 = t_hi * C128;  // < 2^93
    u128 s2 = s + f2;
    carries += (s2 < s);
    s = s2;
    while (carries) {
        u128 add = (u128)carries * C128;
        u128 ns = s + add;
        carries = (ns < s);
        s = ns;
    }
    if (s >= MOD) s -= MOD;
    return s;
}

static inline u128 fexp(u128 b, u128 e) {
    u128 r = 1;
    while (e) {
        if (e & 1) r = fmul(r, b);
        b = fmul(b, b);
        e >>= 1;
    }
    return r;
}
static inline u128 finv(u128 a) { return a == 0 ? 0 : fexp(a, MOD - 2); }  // winter-math: inv(0) = 0

// get_root_of_unity(k) = G^(2^(40-k))
static inline u128 root_of_unity(unsigned log_n) {
    u128 r = TWO_ADIC_ROOT;
    for (unsigned i = log_n; i < TWO_ADICITY; i++) r = fmul(r, r);
    return r;
}

// Montgomery's trick; zeros map to zeros (winter-math `batch_inversion`)
static inline void batch_inverse(u128* v, size_t n) {
    std::vector<u128> pre(n);
    u128 acc = 1;
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (v[i] != 0) acc = fmul(acc, v[i]);
    }
    acc = finv(acc);
    for (size_t i = n; i-- > 0;) {
        if (v[i] == 0) continue;
        u128 t = fmul(acc, pre[i]);
        acc = fmul(acc, v[i]);
        v[i] = t;
    }
}

static inline void store_le(uint8_t* dst, u128 v) { memcpy(dst, &v, 16); }  // x86-64 is little-endian
static inline u128 load_le(const uint8_t* src) {
    u128 v;
    memcpy(&v, src, 16);
    return v;
}

}  // namespace orc
