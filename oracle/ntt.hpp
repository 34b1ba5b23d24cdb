// ORACLE (test infrastructure, not product): polynomial transforms of winter-math 0.9.0 `fft`,
// restated.  Natural-order in, natural-order out.
//   interpolate_poly            : evaluations over <w_n>        -> coefficients
//   interpolate_poly_with_offset: evaluations over o*<w_n>      -> coefficients
//   evaluate_poly_with_offset   : coefficients (n) , blowup B   -> evaluations over o*<w_{nB}>
// Call sites in the reference: prover/src/lib.rs:55-62 (DefaultTraceLde::new), :65-72.
#pragma once
#include "f128.hpp"
#include <vector>

namespace orc {

static inline unsigned ilog2(size_t n) {
    unsigned k = 0;
    while (((size_t)1 << k) < n) k++;
    return k;
}

static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

// tw[k] = w^k, k < n/2
static inline std::vector<u128> power_table(u128 w, size_t count) {
    std::vector<u128> t(count);
    u128 acc = 1;
    for (size_t i = 0; i < count; i++) {
        t[i] = acc;
        acc = fmul(acc, w);
    }
    return t;
}

// in-place DFT: out[j] = sum_m a[m] w^{mj}; `tw` = powers of the primitive n-th root w (n/2 entries)
static inline void ntt_inplace(u128* a, size_t n, const u128* tw) {
    unsigned lg = ilog2(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, lg);
        if (i < j) {
            u128 t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
    for (size_t half = 1; half < n; half <<= 1) {
        size_t step = n / (2 * half);
        for (size_t base = 0; base < n; base += 2 * half) {
            for (size_t k = 0; k < half; k++) {
                u128 u = a[base + k];
                u128 v = fmul(a[base + k + half], tw[k * step]);
                a[base + k] = fadd(u, v);
                a[base + k + half] = fsub(u, v);
            }
        }
    }
}

static inline void forward_ntt(std::vector<u128>& a) {
    size_t n = a.size();
    if (n < 2) return;
    auto tw = power_table(root_of_unity(ilog2(n)), n / 2);
    ntt_inplace(a.data(), n, tw.data());
}

// evaluations over <w_n> -> coefficients
static inline void interpolate_poly(std::vector<u128>& a) {
    size_t n = a.size();
    if (n < 2) return;
    auto tw = power_table(finv(root_of_unity(ilog2(n))), n / 2);
    ntt_inplace(a.data(), n, tw.data());
    u128 ninv = finv((u128)n);
    for (auto& x : a) x = fmul(x, ninv);
}

// evaluations over offset*<w_n> -> coefficients
static inline void interpolate_poly_with_offset(std::vector<u128>& a, u128 offset) {
    interpolate_poly(a);
    u128 oi = finv(offset), acc = 1;
    for (auto& x : a) {
        x = fmul(x, acc);
        acc = fmul(acc, oi);
    }
}

// coefficients (n) -> evaluations over offset*<w_{n*blowup}>, natural order.
// Done coset by coset like winter-math: out[B*j + c] = p((offset*w_L^c) * w_n^j).
static inline std::vector<u128> evaluate_poly_with_offset(const std::vector<u128>& p, u128 offset, size_t blowup) {
    size_t n = p.size(), L = n * blowup;
    std::vector<u128> out(L);
    auto tw = power_table(root_of_unity(ilog2(n)), n / 2 ? n / 2 : 1);
    u128 wl = root_of_unity(ilog2(L));
    std::vector<u128> tmp(n);
    u128 shift = offset;
    for (size_t c = 0; c < blowup; c++) {
        u128 acc = 1;
        for (size_t m = 0; m < n; m++) {
            tmp[m] = fmul(p[m], acc);
            acc = fmul(acc, shift);
        }
        if (n > 1) ntt_inplace(tmp.data(), n, tw.data());
        for (size_t j = 0; j < n; j++) out[blowup * j + c] = tmp[j];
        shift = fmul(shift, wl);
    }
    return out;
}

static inline u128 eval_horner(const u128* p, size_t n, u128 x) {
    u128 acc = 0;
    for (size_t i = n; i-- > 0;) acc = fadd(fmul(acc, x), p[i]);
    return acc;
}

}  // namespace orc
