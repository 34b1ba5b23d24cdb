// Out-of-domain evaluation and DEEP composition (kernels K7, K8 of SURVEY 8a').
// Replaces winter-prover's TracePolyTable::get_ood_frame, CompositionPoly::evaluate_at and
// DeepCompositionPoly::{add_trace_polys, add_composition_poly, evaluate} that
// `Prover::prove` runs between the constraint commitment and FRI (vm/src/lib.rs:26).
//
// Polynomials are kept as "offset-scaled" coefficients a'_m = a_m * 3^m (what the LDE kernels consume), so
// p(x) = sum_m a'_m (x/3)^m.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../common.h"

namespace ezk {

// pw[k][m] = y_k^m, m < n = 2^log_n, k < npoints (npoints <= 2; pitch n)
int power_table(cudaStream_t s, uint32_t log_n, const uint64_t y[2][2], uint32_t npoints, uint4* pw);

// out[(c * npoints + k)] = sum_m coeff[c][m] * y_k^m   for c < ncols, k < npoints (npoints <= 2), as dot products with
// the power table `pw` of power_table() (independent terms instead of a Horner chain).
// scratch: ncols * npoints * blocks_per_col elements, blocks_per_col = max(1, n / 16384).
int eval_polys(cudaStream_t s, const uint4* coeff, uint64_t pitch, uint32_t ncols, uint32_t log_n, const uint4* pw,
               uint32_t npoints, uint4* scratch, uint4* out);

// pq[0][m] = sum_{c<28} tc[c] * tcoeff[c][m];  pq[1][m] = sum_{j<7} cc[j] * ccoeff[j][m]   (pq pitch = n)
int deep_combine_coeffs(cudaStream_t s, const uint4* tcoeff, uint64_t tpitch, const uint4* ccoeff, uint64_t cpitch,
                        uint32_t log_n, const uint4* deep_coeffs /* 28 + 7 */, uint4* pq);

// deep[i] = ((P_i + Q_i - s1)(x_i - zg) + (P_i - s2)(x_i - z)) * inv_den[i],  x_i = 3 w_L^i,
// inv_den[i] = 1/((x_i - z)(x_i - zg));  pq_lde: 2 columns of L (pitch L)
struct DeepScalars {
    uint64_t z[2], zg[2], s1[2], s2[2];
};
int deep_pointwise(cudaStream_t s, const uint4* root_fwd, const uint4* pq_lde, uint32_t log_L, const uint4* inv_den,
                   DeepScalars sc, uint4* deep, RowShard sh = RowShard());  // multi-GPU: packed rows of this rank

// The same DEEP evaluations from the LDE tables instead of the coefficient tables (the pointwise
// formula on LDE rows gives identical values): deep[t] for the packed rows t of this rank, reading only rows it owns.
// Used by the multi-GPU path, where it needs no communication.  tlde: 28 columns, clde: 7 columns (packed row order).
int deep_from_rows(cudaStream_t s, const uint4* root_fwd, const uint4* tlde, uint64_t tpitch, const uint4* clde, uint64_t cpitch,
                   uint32_t log_L, const uint4* deep_coeffs /* 28 + 7 */, const uint4* inv_den, DeepScalars sc, uint4* deep,
                   RowShard sh);

// *flag |= 2 when any of the `count` elements is >= M (the kernels assume canonical input, like BaseElement's memory)
int check_canonical(cudaStream_t s, const uint4* v, uint64_t count, uint32_t* flag);

// *flag |= 1 when any of the `count` elements is non-zero
int check_all_zero(cudaStream_t s, const uint4* v, uint64_t count, uint32_t* flag);

}  // namespace ezk
