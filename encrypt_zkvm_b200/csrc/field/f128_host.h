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

// x * 2^128 = x * (45*2^40 - 1) (mod M): shift-and-subtract form, applied to a 128-bit `hi`
// word of a 256-bit product; the partial result is folded again until it fits in 128 bits.
inline Fp operator*(Fp a, Fp b) {
    const uint64_t a0 = (uint64_t)a.v, a1 = (uint64_t)(a.v >> 64), b0 = (uint64_t)b.v, b1 = (uint64_t)(b.v >> 64);
    const u128 ll = (u128)a0 * b0, lh = (u128)a0 * b1, hl = (u128)a1 * b0, hh = (u128)a1 * b1;
    u128 cross = lh + hl;
    const u128 cross_carry = cross < lh ? ((u128)1 << 64) : 0;
    u128 lo = ll + (cross << 64);
    u128 hi = hh + (cross >> 64) + cross_carry + (lo < ll ? 1 : 0);
    // fold while hi != 0: value = lo + hi*2^128 = lo + (hi*45 << 40) - hi
    while (hi != 0) {
        // hi*45 as 192 bits
        const u128 h_lo = (u128)(uint64_t)hi * 45, h_hi = (u128)(uint64_t)(hi >> 64) * 45;
        const u128 m_lo = h_lo + (h_hi << 64);
        const u128 m_hi = (h_hi >> 64) + (m_lo < h_lo ? 1 : 0);
        // (m << 40) as up to 256 bits: s_hi:s_lo
        const u128 s_lo = m_lo << 40;
        const u128 s_hi = (m_hi << 40) | (m_lo >> 88);
        // lo + s - hi
        u128 nlo = lo + s_lo;
        u128 nhi = s_hi + (nlo < lo ? 1 : 0);
        if (nlo < hi) nhi -= 1;  // borrow (total is non-negative because s >= hi)
        nlo -= hi;
        lo = nlo;
        hi = nhi;
    }
    return Fp::reduce(lo);
}

inline Fp pow(Fp b, u128 e) {
    Fp r(1);
    while (e) {
        if (e & 1) r = r * b;
        b = b * b;
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
