// Batched column NTT / coset-LDE over f128 (kernels K1, K2 of SURVEY 8a').
// Replaces winter-math's fft::interpolate_poly / evaluate_poly_with_offset as used by
// DefaultTraceLde::new (prover/src/lib.rs:55-62), CompositionPoly::new and the DEEP/FRI LDEs.
#pragma once
#include <cstdint>
#include <map>
#include <cuda_runtime.h>

namespace ezk {

struct NttTables {
    // each table: 2 * 2^14 elements (two-level powers, see f128.cuh)
    uint4* root_fwd = nullptr;  // powers of w = primitive 2^28-th root of unity
    uint4* root_inv = nullptr;  // powers of w^-1
    uint4* off_fwd = nullptr;   // powers of the domain offset o = 3
    uint4* off_inv = nullptr;   // powers of o^-1
    // compact per-size tables: entry (1 << k) + e = w_{2^k}^e (forward) / w_{2^k}^-e (inverse), k <= 11
    uint4* tw_fwd = nullptr;
    uint4* tw_inv = nullptr;
    // the same tables in precomputed form (4 x 16 bytes per entry), used by the passes whose kPreTw is set (ntt.cu)
    uint4* twp_fwd = nullptr;
    uint4* twp_inv = nullptr;
    // full inter-pass twiddle tables of the strided passes, built on first use per (size, direction) and kept
    // (16 MiB for a plain 2^20 pass, 128 MiB for the 2^20 LDE); sizes above the limit use the two-level tables
    std::map<uint64_t, uint4*>* big_tables = nullptr;
    uint64_t big_table_limit_bytes = 1ull << 30;
    int max_tile_log = 10;      // largest single-CTA transform (2^10 points x 4 lanes, 2^9 x 8 lanes)
    // output scaling of the trace interpolation as one table, S[m] = c * o^m (c = 1/n), kept for the latest length:
    // one load + one product per coefficient instead of a two-level power lookup + up to three products
    struct ScaleTable {
        uint4* d = nullptr;
        uint32_t log_n = 0;
        uint64_t c[2] = {0, 0};
    };
    ScaleTable* scale_table = nullptr;
};

void ntt_tables_init(NttTables& t);
void ntt_tables_free(NttTables& t);

// Output scaling applied by the last pass: out[m] *= cvec[m >> chunk_shift] * (use_offset ? o^m : 1)
struct NttScale {
    uint64_t cvec[8][2];   // up to 8 constants {lo, hi}; cvec[0] = 1 when unused
    uint32_t chunk_shift;  // 63 => always cvec[0]
    uint32_t use_offset;   // 0 none, 1 multiply by o^m (forward offset table), 2 multiply by o^-m
    uint32_t enabled;
};

// Natural-order DFT of `ncols` columns of n = 2^log_n elements: dst[c][j] = scale * sum_m src[c][m] w^(mj),
// w = primitive n-th root (inverse root when `inverse`). `src` is left intact; `work` must hold ncols * n
// elements when n > 2^max_tile_log (unused otherwise). src/dst pitches in elements. Returns the number of
// kernel launches.
int ntt_columns(const NttTables& t, cudaStream_t s, const uint4* src, uint64_t src_pitch, uint4* dst, uint64_t dst_pitch,
                uint4* work, uint32_t ncols, uint32_t log_n, bool inverse, const NttScale* scale);

// Which of the 8 LDE cosets a prover computes: coset k = base + step * k, k < 1 << count_log.  {3, 0, 1} = all of
// them (single GPU); rank r of G = 2^g ranks owns {r, r + G, ...} = {3 - g, r, G}.
struct CosetSet {
    uint32_t count_log = 3, base = 0, step = 1;
};

// Coset low-degree extension with blowup 8: coeff[c][m] must already be scaled by o^m.  For the k-th coset of `cs`
// (coset number q_k = base + step * k) and j < n:
//     lde[c][(j << count_log) + k] = sum_m coeff[c][m] * (w_L^(q_k))^m * w_n^(mj)  =  column c at LDE row 8 j + q_k,
// i.e. the rows of the chosen cosets in ascending order ("packed" row order; with all 8 cosets it is the natural order
// over the LDE domain o*<w_L>).  lde_pitch >= n << count_log.
// tmp must hold ncols * (number of cosets) * n elements when n > 2^max_tile_log (unused otherwise).
int lde_columns(const NttTables& t, cudaStream_t s, const uint4* coeff, uint64_t coeff_pitch, uint4* lde,
                uint64_t lde_pitch, uint4* tmp, uint32_t ncols, uint32_t log_n, CosetSet cs = CosetSet());

// Multi-GPU composition step: a slice [m0, m0 + count) of the 7 composition columns (offset-scaled coefficients) from
// the same slice of the plain inverse transforms of the constraint evaluations on each of the 8 LDE cosets
// (y[c * count + k], out[j * count + k]; see composition_recombine_kernel).  scale[j] = 3^(-nj) / L.
// *flag |= 1 when the eighth column is not identically zero on the slice.
int composition_recombine(const NttTables& t, cudaStream_t s, const uint4* y, uint32_t log_n, uint64_t m0, uint64_t count,
                          const uint64_t scale[8][2], uint4* out, uint32_t* flag);

}  // namespace ezk
