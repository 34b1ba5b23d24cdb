// OOD evaluation + DEEP composition kernels (K7, K8).  See compose.cuh.
#include "compose.cuh"
#include "../common.h"
#include "../field/f128.cuh"

namespace ezk {

using namespace dev;

namespace {

constexpr int kEvalThreads = 256;
constexpr uint32_t kEvalChunkLog = 14;  // coefficients per block

// pw[k][m] = y_k^m for m < n, k < npoints.  Thread t of a block starts from y^(base + t) (square-and-multiply) and
// walks its residue class with one product per entry.
__global__ void __launch_bounds__(kEvalThreads) power_table_kernel(uint32_t log_n, fe y0, fe y1, uint32_t npoints,
                                                                  uint4* __restrict__ pw) {
    const uint64_t n = 1ull << log_n;
    const uint64_t chunk = n < (1ull << kEvalChunkLog) ? n : (1ull << kEvalChunkLog);
    const uint64_t base = (uint64_t)blockIdx.x * chunk;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    if (t >= chunk) return;
    const fe ys[2] = {y0, y1};
    for (uint32_t k = 0; k < npoints; k++) {
        const fe Y = fe_pow(ys[k], T);
        fe p = fe_pow(ys[k], base + t);
        uint4* dst = pw + (uint64_t)k * n + base;
        for (uint64_t m = t; m < chunk; m += T) {
            fe_store(dst + m, p);
            p = fe_mul(p, Y);
        }
    }
}

// acc[k] += sum over this thread's coefficients of a[m] * pw[k][m]; U coefficients per step, all their loads issued
// before the first product (the kernel is otherwise bound by the latency of its loads)
template <int NP, class AR>
__device__ __forceinline__ bool eval_thread(const uint4* __restrict__ col, const uint4* __restrict__ pw, uint64_t n,
                                            uint64_t chunk, uint32_t t, uint32_t T, fe (&acc)[NP]) {
    AR ar;
#pragma unroll
    for (int k = 0; k < NP; k++) acc[k] = fe_zero();
    constexpr int U = NP == 2 ? 2 : 4;
    uint64_t m = t;
    for (; m + (U - 1) * T < chunk; m += U * T) {
        fe a[U], w[NP][U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            a[u] = fe_ldg(col + m + u * T);
#pragma unroll
            for (int k = 0; k < NP; k++) w[k][u] = fe_ldg(pw + (uint64_t)k * n + m + u * T);
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < NP; k++) acc[k] = ar.add(acc[k], ar.mul(a[u], w[k][u]));
    }
    for (; m < chunk; m += T) {
        const fe a = fe_ldg(col + m);
#pragma unroll
        for (int k = 0; k < NP; k++) acc[k] = ar.add(acc[k], ar.mul(a, fe_ldg(pw + (uint64_t)k * n + m)));
    }
    return ar.tainted();
}

// block (bx, col): partial dot products of coefficients [bx*chunk, (bx+1)*chunk) with the power tables of up to 2
// points.  Every term is independent (no Horner chain), so the products of a thread overlap.
template <int NP>
__global__ void __launch_bounds__(kEvalThreads, 4) eval_partial_kernel(const uint4* __restrict__ coeff, uint64_t pitch,
                                                                      uint32_t log_n, const uint4* __restrict__ pw,
                                                                      uint4* __restrict__ scratch) {
    __shared__ uint4 red[NP][kEvalThreads];
    const uint64_t n = 1ull << log_n;
    const uint64_t chunk = n < (1ull << kEvalChunkLog) ? n : (1ull << kEvalChunkLog);
    const uint64_t base = (uint64_t)blockIdx.x * chunk;
    const uint4* col = coeff + (uint64_t)blockIdx.y * pitch + base;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    fe acc[NP];
    if (eval_thread<NP, Arith<true>>(col, pw + base, n, chunk, t, T, acc))
        eval_thread<NP, Arith<false>>(col, pw + base, n, chunk, t, T, acc);  // a rare tail of the fast arithmetic: redo exactly
#pragma unroll
    for (int k = 0; k < NP; k++) fe_store(&red[k][t], acc[k]);
    __syncthreads();
    for (uint32_t h = T / 2; h >= 1; h >>= 1) {
        if (t < h) {
#pragma unroll
            for (int k = 0; k < NP; k++) fe_store(&red[k][t], fe_add(fe_load(&red[k][t]), fe_load(&red[k][t + h])));
        }
        __syncthreads();
    }
    if (t < NP) scratch[((uint64_t)blockIdx.y * NP + t) * gridDim.x + blockIdx.x] = red[t][0];
}

__global__ void eval_final_kernel(const uint4* __restrict__ scratch, uint32_t nblocks, uint32_t nout,
                                  uint4* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout) return;
    fe acc = fe_zero();
    for (uint32_t b = 0; b < nblocks; b++) acc = fe_add(acc, fe_load(scratch + (uint64_t)t * nblocks + b));
    fe_store(out + t, acc);
}

__global__ void __launch_bounds__(256) deep_combine_kernel(const uint4* __restrict__ tcoeff, uint64_t tpitch,
                                                          const uint4* __restrict__ ccoeff, uint64_t cpitch, uint64_t n,
                                                          const uint4* __restrict__ dc, uint4* __restrict__ pq) {
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    fe p = fe_zero(), q = fe_zero();
#pragma unroll 4
    for (int c = 0; c < 28; c++) p = fe_add(p, fe_mul(fe_ldg(dc + c), fe_ldg(tcoeff + (uint64_t)c * tpitch + m)));
#pragma unroll
    for (int j = 0; j < 7; j++) q = fe_add(q, fe_mul(fe_ldg(dc + 28 + j), fe_ldg(ccoeff + (uint64_t)j * cpitch + m)));
    fe_store(pq + m, p);
    fe_store(pq + n + m, q);
}

__global__ void __launch_bounds__(256) deep_pointwise_kernel(const uint4* __restrict__ roots,
                                                            const uint4* __restrict__ pq_lde, uint32_t log_L,
                                                            const uint4* __restrict__ inv_den, DeepScalars sc,
                                                            RowShard sh, uint4* __restrict__ deep) {
    const uint64_t L = 1ull << log_L;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (L >> sh.world_log)) return;
    const uint64_t i = sh.global_row(t);
    fe w = fe_root_pow(roots, log_L, i);
    fe x = fe_add(fe_add(w, w), w);
    fe P = fe_ldg(pq_lde + i), Q = fe_ldg(pq_lde + L + i);
    fe t1 = fe_sub(fe_add(P, Q), fe_make(sc.s1[0], sc.s1[1]));
    fe t2 = fe_sub(P, fe_make(sc.s2[0], sc.s2[1]));
    fe num = fe_add(fe_mul(t1, fe_sub(x, fe_make(sc.zg[0], sc.zg[1]))), fe_mul(t2, fe_sub(x, fe_make(sc.z[0], sc.z[1]))));
    fe_store(deep + t, fe_mul(num, fe_ldg(inv_den + t)));
}

// DEEP composition straight from the LDE tables (multi-GPU: everything a rank needs is in the rows it owns):
// P_i = sum_c dc[c] T_c(x_i), Q_i = sum_j dc[28 + j] H_j(x_i), then the same quotient as deep_pointwise_kernel.
// Flagged arithmetic with an exact redo of the row, as in the constraint kernel.
template <class AR>
__device__ __forceinline__ bool deep_row(const uint4* __restrict__ roots, const uint4* __restrict__ tlde, uint64_t tpitch,
                                         const uint4* __restrict__ clde, uint64_t cpitch, uint32_t log_L,
                                         const uint4* __restrict__ dc, fe inv, const DeepScalars& sc, uint64_t t, uint64_t i,
                                         fe& out) {
    AR ar;
    fe P = fe_zero(), Q = fe_zero();
#pragma unroll 7
    for (int c = 0; c < 28; c++) P = ar.add(P, ar.mul(fe_ldg(dc + c), fe_ldg(tlde + (uint64_t)c * tpitch + t)));
#pragma unroll
    for (int j = 0; j < 7; j++) Q = ar.add(Q, ar.mul(fe_ldg(dc + 28 + j), fe_ldg(clde + (uint64_t)j * cpitch + t)));
    const fe w = fe_root_pow(roots, log_L, i);
    const fe x = ar.add(ar.add(w, w), w);
    const fe t1 = ar.sub(ar.add(P, Q), fe_make(sc.s1[0], sc.s1[1]));
    const fe t2 = ar.sub(P, fe_make(sc.s2[0], sc.s2[1]));
    const fe num = ar.add(ar.mul(t1, ar.sub(x, fe_make(sc.zg[0], sc.zg[1]))), ar.mul(t2, ar.sub(x, fe_make(sc.z[0], sc.z[1]))));
    out = ar.mul(num, inv);
    return ar.tainted();
}
__global__ void __launch_bounds__(256) deep_rows_kernel(const uint4* __restrict__ roots, const uint4* __restrict__ tlde,
                                                       uint64_t tpitch, const uint4* __restrict__ clde, uint64_t cpitch,
                                                       uint32_t log_L, const uint4* __restrict__ dc,
                                                       const uint4* __restrict__ inv_den, DeepScalars sc, RowShard sh,
                                                       uint4* __restrict__ deep) {
    const uint64_t L = 1ull << log_L;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (L >> sh.world_log)) return;
    const uint64_t i = sh.global_row(t);
    const fe inv = fe_ldg(inv_den + t);
    fe r;
    if (deep_row<Arith<true>>(roots, tlde, tpitch, clde, cpitch, log_L, dc, inv, sc, t, i, r))
        deep_row<Arith<false>>(roots, tlde, tpitch, clde, cpitch, log_L, dc, inv, sc, t, i, r);
    fe_store(deep + t, r);
}

__global__ void all_zero_kernel(const uint4* __restrict__ v, uint64_t count, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    for (; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 x = __ldg(v + i);
        bad |= (x.x | x.y | x.z | x.w) != 0;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// flag |= 2 when an element is not a canonical field element (>= M = 2^128 - 45*2^40 + 1)
__global__ void canonical_kernel(const uint4* __restrict__ v, uint64_t count, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    for (; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 x = __ldg(v + i);
        bad |= (x.w == 0xFFFFFFFFu) && (x.z == 0xFFFFFFFFu) && (x.y > 0xFFFFD300u || (x.y == 0xFFFFD300u && x.x >= 1u));
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 2u);
}

}  // namespace

int power_table(cudaStream_t s, uint32_t log_n, const uint64_t y[2][2], uint32_t npoints, uint4* pw) {
    uint32_t nblocks = log_n > kEvalChunkLog ? 1u << (log_n - kEvalChunkLog) : 1;
    {
        LaunchScope ls(s, K_EVAL_POLYS, ((uint64_t)npoints << log_n) * 16);
        power_table_kernel<<<nblocks, kEvalThreads, 0, s>>>(log_n, fe_make(y[0][0], y[0][1]), fe_make(y[1][0], y[1][1]), npoints, pw);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int eval_polys(cudaStream_t s, const uint4* coeff, uint64_t pitch, uint32_t ncols, uint32_t log_n, const uint4* pw,
               uint32_t npoints, uint4* scratch, uint4* out) {
    uint32_t nblocks = log_n > kEvalChunkLog ? 1u << (log_n - kEvalChunkLog) : 1;
    dim3 grid(nblocks, ncols);
    {
        LaunchScope ls(s, K_EVAL_POLYS, ((uint64_t)(ncols + npoints) << log_n) * 16);
        if (npoints == 2)
            eval_partial_kernel<2><<<grid, kEvalThreads, 0, s>>>(coeff, pitch, log_n, pw, scratch);
        else
            eval_partial_kernel<1><<<grid, kEvalThreads, 0, s>>>(coeff, pitch, log_n, pw, scratch);
    }
    EZK_CUDA(cudaGetLastError());
    uint32_t nout = ncols * npoints;
    {
        LaunchScope ls(s, K_EVAL_POLYS, (uint64_t)nout * nblocks * 16);
        eval_final_kernel<<<(nout + 63) / 64, 64, 0, s>>>(scratch, nblocks, nout, out);
    }
    EZK_CUDA(cudaGetLastError());
    return 2;
}

int deep_combine_coeffs(cudaStream_t s, const uint4* tcoeff, uint64_t tpitch, const uint4* ccoeff, uint64_t cpitch,
                        uint32_t log_n, const uint4* deep_coeffs, uint4* pq) {
    const uint64_t n = 1ull << log_n;
    {
        LaunchScope ls(s, K_DEEP_COMBINE, n * 16 * (35 + 2));
        deep_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tcoeff, tpitch, ccoeff, cpitch, n, deep_coeffs, pq);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int deep_pointwise(cudaStream_t s, const uint4* root_fwd, const uint4* pq_lde, uint32_t log_L, const uint4* inv_den,
                   DeepScalars sc, uint4* deep, RowShard sh) {
    const uint64_t L = (1ull << log_L) >> sh.world_log;
    {
        LaunchScope ls(s, K_DEEP_POINTWISE, L * 16 * 4);
        deep_pointwise_kernel<<<(unsigned)((L + 255) / 256), 256, 0, s>>>(root_fwd, pq_lde, log_L, inv_den, sc, sh, deep);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int deep_from_rows(cudaStream_t s, const uint4* root_fwd, const uint4* tlde, uint64_t tpitch, const uint4* clde, uint64_t cpitch,
                   uint32_t log_L, const uint4* deep_coeffs, const uint4* inv_den, DeepScalars sc, uint4* deep, RowShard sh) {
    const uint64_t rows = (1ull << log_L) >> sh.world_log;
    {
        LaunchScope ls(s, K_DEEP_POINTWISE, rows * 16 * (35 + 2));
        deep_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(root_fwd, tlde, tpitch, clde, cpitch, log_L, deep_coeffs,
                                                                        inv_den, sc, sh, deep);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int check_canonical(cudaStream_t s, const uint4* v, uint64_t count, uint32_t* flag) {
    unsigned blocks = (unsigned)((count + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    {
        LaunchScope ls(s, K_ALL_ZERO, count * 16);
        canonical_kernel<<<blocks, 256, 0, s>>>(v, count, flag);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int check_all_zero(cudaStream_t s, const uint4* v, uint64_t count, uint32_t* flag) {
    unsigned blocks = (unsigned)((count + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    {
        LaunchScope ls(s, K_ALL_ZERO, count * 16);
        all_zero_kernel<<<blocks, 256, 0, s>>>(v, count, flag);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ezk
