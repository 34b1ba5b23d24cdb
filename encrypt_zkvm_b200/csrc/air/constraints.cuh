// Fused constraint evaluation over the LDE domain (kernel K5 of SURVEY 8a').
// Replaces winter-prover's DefaultConstraintEvaluator::evaluate + ConstraintEvaluationTable::combine as
// configured at prover/src/lib.rs:65-72, with `ProcessorAir` (air/src/lib.rs:62-206) inlined.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../common.h"

namespace ezk {

struct ConstraintParams {  // device-resident, rebuilt for every proof (depends on n, coefficients, public inputs)
    uint64_t tcoef[20][2];     // transition composition coefficients
    uint64_t bcoef[22][2];     // boundary composition coefficients, in sorted-assertion order
    uint64_t bval[22][2];      // asserted values (12 zeros for step 0, then 10 values for step n-2)
    uint64_t bsum1[2];         // sum_{k >= 12} bcoef[k] * bval[k]: the constant part of the step n-2 boundary sum
    uint32_t bcol[22];         // asserted columns
    uint32_t delta;            // LWE delta (fhe/src/parameters.rs:17)
    uint64_t inv_zn[8][2];     // 1 / (x^n - 1) for LDE step i = 8j + c (depends on c only)
    uint64_t g_last[2];        // g^(n-2)
    uint64_t g_last2[2];       // g^(n-1)
    uint64_t inv_mds[16][2];   // crypto/src/rescue.rs:216-233
    uint64_t ptable[128 * 9][2];  // periodic values, row = step mod 128
    // the constant multipliers once more in precomputed form (c * 2^(32 i), i = 0..3: fe_pre in f128.cuh), so that the
    // 58 products of a row that have a per-proof constant operand cost 16 IMAD.WIDE + one small fold each
    uint64_t tcoef_pre[20][4][2];
    uint64_t bcoef_pre[22][4][2];
    uint64_t inv_mds_pre[16][4][2];
};

// out[i] = 1 / ((x_i - a)(x_i - b)),  x_i = 3 * w_L^i,  i < L = 2^log_L
// (multi-GPU: out[t] for the packed rows t of this rank, x taken at global_row(t))
int domain_pair_inverse(cudaStream_t s, const uint4* root_fwd, const uint4* root_inv, uint32_t log_L, const uint64_t a[2],
                        const uint64_t b[2], uint4* out, RowShard sh = RowShard());

// combined[i] = T_i / z_t(x_i) + B0_i / (x_i - 1) + B1_i / (x_i - g^(n-2))   
// lde: column-major 28 x (L / world) rows in packed order; inv_den[t] = 1/((x_i - 1)(x_i - g^(n-2))), i = global_row(t)
// coset_major: combined[k * n + j] for the k-th coset this rank owns (LDE row 8 j + coset) instead of packed row order
int evaluate_constraints(cudaStream_t s, const uint4* root_fwd, const uint4* lde, uint64_t pitch, uint32_t log_L,
                         const ConstraintParams* params, const uint4* inv_den, uint4* combined, RowShard sh = RowShard(),
                         bool coset_major = false);

// parity helper: 20 transition values for explicit frames (cur/nxt: nframes x 28, periodic: nframes x 9, out: nframes x 20)
int evaluate_frames(cudaStream_t s, const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                    const ConstraintParams* params, uint4* out);

// parity helper for the production (selector-grouped, flagged) path: out[f] = sum_j params->tcoef[j] * r_j(frame f)
int evaluate_frames_sum(cudaStream_t s, const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                        const ConstraintParams* params, uint4* out);

}  // namespace ezk
