// Host-side pieces of `ProcessorAir` that are data, not arithmetic over the trace: the 9 periodic columns
// (air/src/lib.rs:201-225: CYCLE_MASK + the 8 ARK columns of crypto/src/rescue.rs:120-134,235-378) as
// polynomials over <w_16>.  Used by the prover (periodic table over the LDE domain) and the verifier (values at
// the out-of-domain point).
#pragma once
#include "../field/f128_host.h"
#include "../../../include/ezkvm_rescue_constants.h"
#include <vector>

namespace ezk {

struct RescuePair64 {
    uint64_t lo, hi;
};
inline Fp rescue_const(const RescuePair64& p) { return Fp(((u128)p.hi << 64) | p.lo); }
inline const RescuePair64* rescue_inv_mds() {
    static const RescuePair64 k[16] = {EZK_RESCUE_INV_MDS_INIT};
    return k;
}
inline const RescuePair64* rescue_ark() {
    static const RescuePair64 k[128] = {EZK_RESCUE_ARK_INIT};
    return k;
}

// interpolate 16 values over <w_16> (naive inverse DFT; 9 columns x 256 products per proof)
inline std::vector<Fp> interpolate16(const Fp* values) {
    const Fp winv = inverse(root_of_unity(4)), ninv = inverse(Fp::from_u64(16));
    std::vector<Fp> coeffs(16);
    for (int k = 0; k < 16; k++) {
        Fp acc, wk = pow(winv, k), p(1);
        for (int i = 0; i < 16; i++) {
            acc = acc + values[i] * p;
            p = p * wk;
        }
        coeffs[k] = acc * ninv;
    }
    return coeffs;
}

inline Fp horner(const std::vector<Fp>& p, Fp x) {
    Fp acc;
    for (size_t i = p.size(); i-- > 0;) acc = acc * x + p[i];
    return acc;
}

// coefficient vectors of the 9 periodic columns: [mask (14 ones, 2 zeros), ark_0 .. ark_7]
inline std::vector<std::vector<Fp>> periodic_polys() {
    std::vector<std::vector<Fp>> polys;
    for (uint32_t p = 0; p < 9; p++) {
        Fp vals[16];
        for (uint32_t i = 0; i < 16; i++) vals[i] = p == 0 ? Fp::from_u64(i < 14 ? 1 : 0) : rescue_const(rescue_ark()[i * 8 + (p - 1)]);
        polys.push_back(interpolate16(vals));
    }
    return polys;
}

}  // namespace ezk
