// FRI fold + remainder kernels (K9).  See fri.cuh.
#include "fri.cuh"
#include "../common.h"
#include "../field/f128.cuh"

namespace ezk {

using namespace dev;

namespace {

__device__ __forceinline__ fe ld2(const uint64_t v[2]) { return fe_make(v[0], v[1]); }

// One thread per folded position: 8 coalesced loads (one per j), 8-point inverse DFT in registers
// (decimation in frequency, outputs land bit-reversed), Horner in beta = alpha / x_i.
__global__ void __launch_bounds__(256) fri_fold_kernel(const uint4* __restrict__ root_inv,
                                                      const uint4* __restrict__ evals, uint32_t log_s, FriFoldConsts c,
                                                      RowShard sh, uint4* __restrict__ next) {
    // multi-GPU: `evals` and `next` hold this rank's positions in ascending order (position p = world * t + rank
    // lives at t), so a row is again 8 values at stride m and only the domain point needs the global index
    const uint64_t m = (1ull << (log_s - 3)) >> sh.world_log;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    fe a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = fe_ldg(evals + i + (uint64_t)j * m);
    const fe w1 = ld2(c.zinv[1]), w2 = ld2(c.zinv[2]), w3 = ld2(c.zinv[3]);
    // stage 1 (span 4)
    fe b[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        b[k] = fe_add(a[k], a[k + 4]);
        b[k + 4] = fe_sub(a[k], a[k + 4]);
    }
    b[5] = fe_mul(b[5], w1), b[6] = fe_mul(b[6], w2), b[7] = fe_mul(b[7], w3);
    // stage 2 (span 2)
    fe d[8];
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        d[h] = fe_add(b[h], b[h + 2]);
        d[h + 1] = fe_add(b[h + 1], b[h + 3]);
        d[h + 2] = fe_sub(b[h], b[h + 2]);
        d[h + 3] = fe_mul(fe_sub(b[h + 1], b[h + 3]), w2);
    }
    // stage 3 (span 1): position p holds coefficient bitrev3(p)
    fe e[8];
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
        e[h] = fe_add(d[h], d[h + 1]);
        e[h + 1] = fe_sub(d[h], d[h + 1]);
    }
    // natural order: c_0 = e0, c_1 = e4, c_2 = e2, c_3 = e6, c_4 = e1, c_5 = e5, c_6 = e3, c_7 = e7
    // beta = alpha * x_i^-1 = (alpha/3) * w_s^-i
    const fe beta = fe_mul(ld2(c.alpha_oinv), fe_root_pow(root_inv, log_s, sh.global_row(i)));
    fe acc = e[7];
    acc = fe_add(fe_mul(acc, beta), e[3]);
    acc = fe_add(fe_mul(acc, beta), e[5]);
    acc = fe_add(fe_mul(acc, beta), e[1]);
    acc = fe_add(fe_mul(acc, beta), e[6]);
    acc = fe_add(fe_mul(acc, beta), e[2]);
    acc = fe_add(fe_mul(acc, beta), e[4]);
    acc = fe_add(fe_mul(acc, beta), e[0]);
    fe_store(next + i, fe_mul(acc, ld2(c.inv8)));
}

// s <= 4096 points: one thread per output coefficient, direct O(s^2) sum (the FRI tail is tiny)
__global__ void fri_remainder_kernel(const uint4* __restrict__ root_inv, const uint4* __restrict__ off_inv,
                                     const uint4* __restrict__ evals, uint32_t log_s, fe inv_s,
                                     uint4* __restrict__ coeffs) {
    const uint32_t s = 1u << log_s;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s) return;
    const fe wk = fe_root_pow(root_inv, log_s, k);  // w_s^-k
    fe acc = fe_zero(), p = fe_one();
    for (uint32_t i = 0; i < s; i++) {
        acc = fe_add(acc, fe_mul(fe_ldg(evals + i), p));
        p = fe_mul(p, wk);
    }
    acc = fe_mul(acc, fe_mul(inv_s, fe_tab_pow(off_inv, k)));
    fe_store(coeffs + k, acc);
}

}  // namespace

int fri_fold(cudaStream_t s, const uint4* root_inv, const uint4* evals, uint32_t log_s, FriFoldConsts c, uint4* next, RowShard sh) {
    const uint64_t m = (1ull << (log_s - 3)) >> sh.world_log;
    {
        LaunchScope ls(s, K_FRI_FOLD, m * 16 * 9);
        fri_fold_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(root_inv, evals, log_s, c, sh, next);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int fri_remainder(cudaStream_t s, const uint4* root_inv, const uint4* off_inv, const uint4* evals, uint32_t log_s,
                  const uint64_t inv_s[2], uint4* coeffs) {
    const uint32_t n = 1u << log_s;
    {
        LaunchScope ls(s, K_FRI_REMAINDER, (uint64_t)n * 32);
        fri_remainder_kernel<<<(n + 63) / 64, 64, 0, s>>>(root_inv, off_inv, evals, log_s, fe_make(inv_s[0], inv_s[1]), coeffs);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ezk
