// Host-side arithmetic in the reference's base field f128 (M = 2^128 - 45*2^40 + 1), used by the
// host half of the prover (VM trace builder, Fiat-Shamir transcript, per-proof scalar set-up).
// Reference: `winterfell::math::fields::f128::BaseElement` as configured at prover/src/lib.rs:4,41.
// Values are canonical u128 (< M), serialized as 16 little-endian bytes.
#pragma once
#include <cstdint>
#include <cstring>

namespace ezk {

typedef unsigned __int128 u128;

struct Fp {
    u128 v;
    Fp() : v(0) {}
    constexpr explicit Fp(u128 x) : v(x) {}
    static constexpr u128 modulus() { return (~(u128)0) - (((u128)45) << 40) + 2; }  // 2^128 - 45*2^40 + 1
    static Fp from_u64(uint64_t x) { return Fp((u128)x); }
    static Fp reduce(u128 x) { return Fp(x >= modulus() ? x - modulus() : x); }
    bool operator==(const Fp& o) const { return v == o.v; }
    bool operator!=(const Fp& o) const { return v != o.v; }
    bool is_zero() const { return v == 0; }
};

inline Fp operator+(Fp a, Fp b) {
    u128 s = a.v + b.v;
    bool wrap = s < a.v;
    if (wrap || s >= Fp::modulus()) s -= Fp::modulus();
    return Fp(s);
}
inline Fp operator-(Fp a, Fp b) { return Fp(a.v >= b.v ? a.v - b.v : a.v + (Fp::modulus() - b.v)); }
inline Fp operator-(Fp a) { return Fp(a.v ? Fp::modulus() - a.v : 0); }

// Multiplication: schoolbook 2x2 product of 64-bit limbs, then two folds with 2^128 = 45*2^40 - 1 (mod M).
//   value = L + H*2^128 = L + (H*45 << 40) - H          (H*45 << 40 >= H, so the difference never borrows out)
// After the first fold the part above 2^128 is < 2^47, after the second it is 0 or 1.  Branches are on
// events of probability ~2^-35 (carry of the second fold) and on the final canonical subtraction only, so
// independent products interleave in the out-of-order core (the VM's Rescue sponge runs four at a time).
namespace detail {
inline Fp fold256(uint64_t r0, uint64_t r1, uint64_t r2, uint64_t r3) {
    const u128 m0 = (u128)r2 * 45, m1 = (u128)r3 * 45 + (uint64_t)(m0 >> 64);
    const uint64_t h0 = (uint64_t)m0, h1 = (uint64_t)m1, h2 = (uint64_t)(m1 >> 64);
    const uint64_t s0 = h0 << 40, s1 = (h1 << 40) | (h0 >> 24), s2 = (h2 << 40) | (h1 >> 24);
    u128 acc = (u128)r0 + s0;
    const uint64_t t0 = (uint64_t)acc;
    acc = (acc >> 64) + r1 + s1;
    const uint64_t t1 = (uint64_t)acc;
    uint64_t t2 = s2 + (uint64_t)(acc >> 64);
    const u128 x = ((u128)t1 << 64) | t0, h = ((u128)r3 << 64) | r2;
    const u128 y = x - h;
    t2 -= (x < h);                                       // < 2^47
    const u128 add = (((u128)t2 * 45) << 40) - t2;       // t2 * (45*2^40 - 1) < 2^93
    u128 z = y + add;
    if (z < y) z += (((u128)45) << 40) - 1;              // wrapped past 2^128: z is small, one more fold
    return Fp::reduce(z);
}
}  // namespace detail

inline Fp operator*(Fp a, Fp b) {
    const uint64_t a0 = (uint64_t)a.v, a1 = (uint64_t)(a.v >> 64), b0 = (uint64_t)b.v, b1 = (uint64_t)(b.v >> 64);
    const u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 t = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    const uint64_t r1 = (uint64_t)t;
    t = (t >> 64) + (p01 >> 64) + (p10 >> 64) + (uint64_t)p11;
    return detail::fold256((uint64_t)p00, r1, (uint64_t)t, (uint64_t)(p11 >> 64) + (uint64_t)(t >> 64));
}

inline Fp square(Fp a) {
    const uint64_t a0 = (uint64_t)a.v, a1 = (uint64_t)(a.v >> 64);
    const u128 p00 = (u128)a0 * a0, p01 = (u128)a0 * a1, p11 = (u128)a1 * a1;
    u128 t = (p00 >> 64) + 2 * (u128)(uint64_t)p01;
    const uint64_t r1 = (uint64_t)t;
    t = (t >> 64) + 2 * (p01 >> 64) + (uint64_t)p11;
    return detail::fold256((uint64_t)p00, r1, (uint64_t)t, (uint64_t)(p11 >> 64) + (uint64_t)(t >> 64));
}

inline Fp pow(Fp b, u128 e) {
    Fp r(1);
    while (e) {
        if (e & 1) r = r * b;
        b = square(b);
        e >>= 1;
    }
    return r;
}
inline Fp inverse(Fp a) { return a.is_zero() ? a : pow(a, Fp::modulus() - 2); }

// two-adic root of unity of order 2^40 = 3^((M-1)/2^40); get_root_of_unity(k) = G^(2^(40-k))
inline Fp root_of_unity(unsigned log_n) {
    Fp r((((u128)0x120532E7B364080AULL) << 64) | 0x86B8723E1920F4AAULL);
    for (unsigned i = log_n; i < 40; i++) r = r * r;
    return r;
}
static constexpr uint64_t kDomainOffset = 3;  // f128 GENERATOR = ProofOptions::domain_offset

inline void fp_store(uint8_t* dst, Fp x) { memcpy(dst, &x.v, 16); }
inline Fp fp_load(const uint8_t* src) {
    Fp x;
    memcpy(&x.v, src, 16);
    return x;
}

}  // namespace ezk
