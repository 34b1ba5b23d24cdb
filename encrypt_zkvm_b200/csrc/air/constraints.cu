// Fused constraint evaluation kernel (K5) and the divisor-inverse kernel.  See constraints.cuh.
#include "constraints.cuh"
#include "processor_air.cuh"
#include "../common.h"

namespace ezk {

using namespace dev;

namespace {

constexpr int kBatch = 64;  // elements per thread in Montgomery batch inversion

__device__ __forceinline__ fe ld2(const uint64_t v[2]) { return fe_make(v[0], v[1]); }

// x_i = 3 * w_L^i
__device__ __forceinline__ fe domain_point(const uint4* __restrict__ roots, uint32_t log_L, uint64_t i) {
    fe w = fe_root_pow(roots, log_L, i);
    return fe_add(fe_add(w, w), w);
}

template <class A>
__device__ __forceinline__ fe root_pow_policy(A& ar, const uint4* __restrict__ tab, uint32_t log_n, uint64_t e) {
    const uint32_t ee = (uint32_t)(e & ((1ull << log_n) - 1)) << (EZK_ROOT_LOG - log_n);
    const uint32_t lo = ee & (EZK_TAB_SIZE - 1), hi = ee >> EZK_TAB_BITS;
    const fe a = fe_ldg(tab + lo), b = fe_ldg(tab + EZK_TAB_SIZE + hi);
    if (hi == 0) return a;
    if (lo == 0) return b;
    return ar.mul(a, b);
}

// out[u] = 1 / ((x - a)(x - b)) at x = 3 w_L^(global_row(u)), u < L / world.  Montgomery batch inversion with
// kBatch elements per thread: the running prefix products are parked in `out` itself (coalesced, re-read on the way
// back), the domain point advances by one multiplication per element, and the a^(M-2) exponentiation (≈ 250
// products) is amortised over kBatch = 64 elements: ≈ 11 products per element.
__global__ void __launch_bounds__(128) pair_inverse_kernel(const uint4* __restrict__ roots, const uint4* __restrict__ roots_inv,
                                                           uint32_t log_L, RowShard sh, fe a, fe b, uint4* __restrict__ out) {
    const uint64_t L = (1ull << log_L) >> sh.world_log;  // rows of this rank, packed index
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;  // a multiple of 8, so global rows advance uniformly
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gstep = nthreads << sh.world_log;             // global_row(u + nthreads) - global_row(u)
    const fe step = fe_root_pow(roots, log_L, gstep), step_inv = fe_root_pow(roots_inv, log_L, gstep);
    fe x = domain_point(roots, log_L, sh.global_row(t));
    fe acc = fe_one();
    for (int q = 0; q < kBatch; q++) {
        const uint64_t u = t + q * nthreads;
        if (u < L) {
            fe_store(out + u, acc);
            acc = fe_mul(acc, fe_mul(fe_sub(x, a), fe_sub(x, b)));
        }
        x = fe_mul(x, step);
    }
    acc = fe_inv(acc);
    for (int q = kBatch - 1; q >= 0; q--) {
        const uint64_t u = t + q * nthreads;
        x = fe_mul(x, step_inv);
        if (u < L) {
            const fe pre = fe_load(out + u);
            fe_store(out + u, fe_mul(acc, pre));
            acc = fe_mul(acc, fe_mul(fe_sub(x, a), fe_sub(x, b)));
        }
    }
}

struct LdeFrame {
    const uint4* __restrict__ lde;
    uint64_t pitch, i, inext;
    __device__ __forceinline__ fe cur(int c) const { return fe_ldg(lde + (uint64_t)c * pitch + i); }
    __device__ __forceinline__ fe nxt(int c) const { return fe_ldg(lde + (uint64_t)c * pitch + inext); }
};

struct SumSink {
    static constexpr bool kGrouped = true;  // only sum_j tcoef_j * r_j is wanted: see eval_transition
    const ConstraintParams* __restrict__ p;
    fe acc;
    template <class A>
    __device__ __forceinline__ fe scaled(A& ar, int j, fe v) { return ar.mul_pre(v, ld_pre(p->tcoef_pre[j])); }
    template <class A>
    __device__ __forceinline__ void add(A& ar, fe v) { acc = ar.add(acc, v); }
    template <class A>
    __device__ __forceinline__ void put(A& ar, int j, fe v) { acc = ar.add(acc, scaled(ar, j, v)); }
};

// combined evaluation of one LDE row; returns true when the FAST arithmetic hit a rare tail (value unusable)
template <class AR>
__device__ __forceinline__ bool constraint_row(const uint4* __restrict__ roots, const uint4* __restrict__ lde, uint64_t pitch,
                                               uint32_t log_L, const ConstraintParams* __restrict__ p,
                                               const uint4* __restrict__ inv_den, RowShard sh, uint64_t packed, uint64_t i,
                                               fe& result) {
    const uint64_t L_local = (1ull << log_L) >> sh.world_log;
    AR ar;
    // the table holds this rank's rows in packed order: row i + 8 sits 8 / world entries after row i
    LdeFrame f{lde, pitch, packed, (packed + sh.frame_step()) & (L_local - 1)};
    fe periodic[9];
    {
        const uint64_t(*row)[2] = &p->ptable[(i & 127) * 9];
#pragma unroll
        for (int k = 0; k < 9; k++) periodic[k] = ld2(row[k]);
    }
    SumSink sink{p, fe_zero()};
    eval_transition(ar, f, periodic, reinterpret_cast<const AirConsts*>(p->inv_mds_pre), p->delta, sink);
    ar.checkpoint();
    fe x = root_pow_policy(ar, roots, log_L, i);
    x = ar.add(ar.add(x, x), x);  // x_i = 3 * w_L^i
    const fe a = ld2(p->g_last);
    // transition part: T * (x - g^(n-2)) (x - g^(n-1)) / (x^n - 1)
    fe t = ar.mul(sink.acc, ar.mul(ar.sub(x, a), ar.sub(x, ld2(p->g_last2))));
    t = ar.mul(t, ld2(p->inv_zn[i & 7]));
    // boundary groups: step 0 (12 assertions, all values zero) and step n-2 (10 assertions)
    fe s0 = fe_zero(), s1 = fe_zero();
#pragma unroll
    for (int k = 0; k < 12; k++) s0 = ar.add(s0, ar.mul_pre(f.cur(p->bcol[k]), ld_pre(p->bcoef_pre[k])));
    ar.checkpoint();
#pragma unroll
    for (int k = 12; k < 22; k++) s1 = ar.add(s1, ar.mul_pre(f.cur(p->bcol[k]), ld_pre(p->bcoef_pre[k])));
    s1 = ar.sub(s1, ld2(p->bsum1));  // sum bcoef_k (cur_k - val_k) = sum bcoef_k cur_k - sum bcoef_k val_k
    ar.checkpoint();
    // B0/(x-1) + B1/(x-a) = (B0 (x-a) + B1 (x-1)) / ((x-1)(x-a))
    const fe num = ar.add(ar.mul(s0, ar.sub(x, a)), ar.mul(s1, ar.sub(x, fe_one())));
    const fe bsum = ar.mul(num, fe_ldg(inv_den + packed));
    result = ar.add(t, bsum);
    return ar.tainted();
}

__device__ __noinline__ fe constraint_row_exact(const uint4* __restrict__ roots, const uint4* __restrict__ lde, uint64_t pitch,
                                                uint32_t log_L, const ConstraintParams* __restrict__ p,
                                                const uint4* __restrict__ inv_den, RowShard sh, uint64_t t, uint64_t i) {
    fe r;
    constraint_row<Arith<false>>(roots, lde, pitch, log_L, p, inv_den, sh, t, i, r);
    return r;
}

template <int THREADS, int MINB, class AR>
__global__ void __launch_bounds__(THREADS, MINB) constraint_kernel(const uint4* __restrict__ roots, const uint4* __restrict__ lde,
                                                                   uint64_t pitch, uint32_t log_L,
                                                                   const ConstraintParams* __restrict__ p,
                                                                   const uint4* __restrict__ inv_den, RowShard sh,
                                                                   uint32_t coset_major, uint4* __restrict__ combined) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ((1ull << log_L) >> sh.world_log)) return;
    const uint64_t i = sh.global_row(t);
    fe r;
    if (constraint_row<AR>(roots, lde, pitch, log_L, p, inv_den, sh, t, i, r))
        r = constraint_row_exact(roots, lde, pitch, log_L, p, inv_den, sh, t, i);
    // packed order (t = rows of this rank in ascending order) or one contiguous array of n values per owned coset
    // (what the per-coset interpolation of the sharded composition step reads)
    const uint32_t cn_log = 3 - sh.world_log;
    const uint64_t at = coset_major ? ((t & ((1u << cn_log) - 1)) << (log_L - 3)) + (t >> cn_log) : t;
    fe_store(combined + at, r);
}

struct ArrayFrame {
    const uint4* c;
    const uint4* n;
    __device__ __forceinline__ fe cur(int k) const { return fe_load(c + k); }
    __device__ __forceinline__ fe nxt(int k) const { return fe_load(n + k); }
};
struct StoreSink {
    static constexpr bool kGrouped = false;  // every constraint value on its own
    uint4* out;
    template <class A>
    __device__ __forceinline__ fe scaled(A&, int, fe v) { return v; }
    template <class A>
    __device__ __forceinline__ void add(A&, fe) {}
    template <class A>
    __device__ __forceinline__ void put(A&, int j, fe v) { fe_store(out + j, v); }
};

__global__ void frames_kernel(const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                              const ConstraintParams* __restrict__ p, uint4* out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nframes) return;
    ArrayFrame f{cur + 28 * t, nxt + 28 * t};
    fe per[9];
    for (int k = 0; k < 9; k++) per[k] = fe_load(periodic + 9 * t + k);
    StoreSink sink{out + 20 * t};
    Arith<false> ar;
    eval_transition(ar, f, per, reinterpret_cast<const AirConsts*>(p->inv_mds_pre), p->delta, sink);
}

// the production path of the constraint kernel at frame level: selector-grouped accumulation of sum_j tcoef_j * r_j
// with the flagged arithmetic and its exact redo (one value per frame)
__global__ void frames_sum_kernel(const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                                  const ConstraintParams* __restrict__ p, uint4* out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nframes) return;
    ArrayFrame f{cur + 28 * t, nxt + 28 * t};
    fe per[9];
    for (int k = 0; k < 9; k++) per[k] = fe_load(periodic + 9 * t + k);
    SumSink sink{p, fe_zero()};
    Arith<true> ar;
    eval_transition(ar, f, per, reinterpret_cast<const AirConsts*>(p->inv_mds_pre), p->delta, sink);
    if (ar.tainted()) {
        Arith<false> exact;
        sink.acc = fe_zero();
        eval_transition(exact, f, per, reinterpret_cast<const AirConsts*>(p->inv_mds_pre), p->delta, sink);
    }
    fe_store(out + t, sink.acc);
}

}  // namespace

int domain_pair_inverse(cudaStream_t s, const uint4* root_fwd, const uint4* root_inv, uint32_t log_L, const uint64_t a[2],
                        const uint64_t b[2], uint4* out, RowShard sh) {
    const uint64_t L = (1ull << log_L) >> sh.world_log;
    const unsigned threads = 128;
    uint64_t need = (L + kBatch - 1) / kBatch;
    unsigned blocks = (unsigned)((need + threads - 1) / threads);
    {
        LaunchScope ls(s, K_PAIR_INVERSE, L * 16);
        pair_inverse_kernel<<<blocks, threads, 0, s>>>(root_fwd, root_inv, log_L, sh, fe_make(a[0], a[1]), fe_make(b[0], b[1]), out);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int evaluate_constraints(cudaStream_t s, const uint4* root_fwd, const uint4* lde, uint64_t pitch, uint32_t log_L,
                         const ConstraintParams* params, const uint4* inv_den, uint4* combined, RowShard sh, bool coset_major) {
    const uint64_t L = (1ull << log_L) >> sh.world_log;
    static int variant = -1;
    if (variant < 0) {
        const char* env = getenv("EZK_CONSTRAINT_VARIANT");
        variant = env ? atoi(env) : 2;  // 256-thread lockstep CTAs, 2 per SM: 5.4 ms at 2^20 against 6.25 without barriers
    }
    {
        LaunchScope ls(s, K_CONSTRAINTS, L * 16 * (28 + 2));  // 28 columns + inv_den read, 1 column written
        // lockstep variants need full CTAs (every thread reaches every barrier): L is a multiple of 512 for n >= 64
        if (variant == 1 && L % 512 == 0)
            constraint_kernel<512, 1, ArithLockstep><<<(unsigned)(L / 512), 512, 0, s>>>(root_fwd, lde, pitch, log_L, params, inv_den, sh, coset_major ? 1u : 0u, combined);
        else if (variant == 2 && L % 256 == 0)
            constraint_kernel<256, 2, ArithLockstep><<<(unsigned)(L / 256), 256, 0, s>>>(root_fwd, lde, pitch, log_L, params, inv_den, sh, coset_major ? 1u : 0u, combined);
        else if (variant == 3 && L % 1024 == 0)
            constraint_kernel<1024, 1, ArithLockstep><<<(unsigned)(L / 1024), 1024, 0, s>>>(root_fwd, lde, pitch, log_L, params, inv_den, sh, coset_major ? 1u : 0u, combined);
        else
            constraint_kernel<128, 4, Arith<true>><<<(unsigned)((L + 127) / 128), 128, 0, s>>>(root_fwd, lde, pitch, log_L, params, inv_den, sh, coset_major ? 1u : 0u, combined);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int evaluate_frames_sum(cudaStream_t s, const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                        const ConstraintParams* params, uint4* out) {
    {
        LaunchScope ls(s, K_FRAMES, (uint64_t)nframes * 16 * (28 * 2 + 9 + 1));
        frames_sum_kernel<<<(nframes + 63) / 64, 64, 0, s>>>(cur, nxt, periodic, nframes, params, out);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int evaluate_frames(cudaStream_t s, const uint4* cur, const uint4* nxt, const uint4* periodic, uint32_t nframes,
                    const ConstraintParams* params, uint4* out) {
    {
        LaunchScope ls(s, K_FRAMES, (uint64_t)nframes * 16 * (28 * 2 + 9 + 20));
        frames_kernel<<<(nframes + 63) / 64, 64, 0, s>>>(cur, nxt, periodic, nframes, params, out);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ezk
