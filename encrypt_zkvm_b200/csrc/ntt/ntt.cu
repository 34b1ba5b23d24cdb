// Batched column NTT / coset-LDE kernels (K1, K2).  See ntt.cuh.
//
// Decomposition (mixed-radix, natural order in and out, no bit-reversal pass over HBM):
//   n = n_1 * n_2 (* n_3);  input index m = m_1 + n_1 m_2 + n_1 n_2 m_3.
//   pass i (i = p..2, "strided"): size-n_i transform over m_i at stride n_1..n_{i-1}, in place, followed by
//       the twiddle w_{N_i}^{lo * j_i}, N_i = n_1..n_i, lo = index below digit i.  A CTA stages n_i x 8
//       elements (8 adjacent lo values = 128-byte segments) in shared memory.
//   pass 1 ("final"): size-n_1 transform over contiguous runs, written out of place to the natural output
//       index j = j_p + n_p (j_{p-1} + ... n_2 j_1); a CTA handles 8 runs whose outputs are adjacent
//       (or the 8 LDE cosets of one run), so global writes are again 128-byte segments.
// Inside a tile: radix-2 decimation-in-frequency stages on shared memory, output taken bit-reversed.
#include "ntt.cuh"
#include "../field/f128.cuh"
#include "../field/f128_host.h"
#include "../common.h"
#include <vector>

namespace ezk {

using namespace dev;

namespace {

constexpr int kLanes = 8;        // adjacent elements handled together (8 x 16 B = one 128-byte line)
constexpr int kPad = kLanes + 1; // shared-memory row pitch (elements) -> conflict-free transposed access
constexpr int kThreads = 512;

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, uint32_t bits) { return __brev(x) >> (32 - bits); }

// in-tile DIF transform over the row index of tile[row * pitch + lane]; rows = 2^log_s
__device__ __forceinline__ void tile_dif(uint4* tile, uint32_t log_s, uint32_t lanes_log, uint32_t pitch,
                                         const uint4* __restrict__ roots) {
    const uint32_t half_total = (1u << (log_s - 1)) << lanes_log;
    for (int lh = (int)log_s - 1; lh >= 0; lh--) {
        const uint32_t h = 1u << lh;
        for (uint32_t b = threadIdx.x; b < half_total; b += blockDim.x) {
            uint32_t lane = b & ((1u << lanes_log) - 1), k = b >> lanes_log;
            uint32_t off = k & (h - 1), blk = k >> lh;
            uint32_t i0 = (blk << (lh + 1)) + off, i1 = i0 + h;
            fe a = fe_load(tile + i0 * pitch + lane), c = fe_load(tile + i1 * pitch + lane);
            fe sum = fe_add(a, c), dif = fe_sub(a, c);
            if (off != 0) {
                // w_S^(off * S / 2h): exponent is a multiple of 2^(28-log_s) -> single table load
                fe w = fe_root_pow(roots, log_s, (uint64_t)off << (log_s - 1 - lh));
                dif = fe_mul(dif, w);
            }
            fe_store(tile + i0 * pitch + lane, sum);
            fe_store(tile + i1 * pitch + lane, dif);
        }
        __syncthreads();
    }
}

struct StridedArgs {
    const uint4* src;
    uint4* dst;
    uint64_t src_pitch, dst_pitch;  // per grid.y column, in elements
    uint32_t log_stride, log_s;     // stride = lo range, S = n_i
    uint32_t coset_first;           // LDE first pass: column y reads coefficient column y/8, scaled by w_L^(c*m), c = y%8
    uint32_t log_L;
    const uint4* roots;
};

__global__ void __launch_bounds__(kThreads) ntt_strided_pass(StridedArgs a) {
    extern __shared__ uint4 tile[];
    const uint32_t S = 1u << a.log_s;
    const uint64_t stride = 1ull << a.log_stride;
    const uint32_t tiles_per_hi = (uint32_t)(stride / kLanes);
    const uint32_t lo0 = (blockIdx.x % tiles_per_hi) * kLanes;
    const uint64_t hi = blockIdx.x / tiles_per_hi;
    const uint64_t base = lo0 + hi * (stride << a.log_s);
    uint32_t col = blockIdx.y, coset = 0;
    const uint4* src;
    if (a.coset_first) {
        coset = col & 7;
        src = a.src + (uint64_t)(col >> 3) * a.src_pitch;
    } else {
        src = a.src + (uint64_t)col * a.src_pitch;
    }
    uint4* dst = a.dst + (uint64_t)col * a.dst_pitch;
    const uint32_t total = S * kLanes;
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        uint32_t lane = e & (kLanes - 1), m = e >> 3;
        uint64_t idx = base + lane + stride * m;
        fe v = fe_load(src + idx);
        if (a.coset_first && coset != 0) v = fe_mul(v, fe_root_pow(a.roots, a.log_L, (uint64_t)coset * idx));
        fe_store(tile + m * kLanes + lane, v);
    }
    __syncthreads();
    tile_dif(tile, a.log_s, 3, kLanes, a.roots);
    const uint32_t log_N = a.log_stride + a.log_s;
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        uint32_t lane = e & (kLanes - 1), pos = e >> 3;
        uint32_t j = bitrev32(pos, a.log_s);
        fe v = fe_load(tile + pos * kLanes + lane);
        uint64_t ex = (uint64_t)(lo0 + lane) * j;
        if (ex != 0) v = fe_mul(v, fe_root_pow(a.roots, log_N, ex));
        fe_store(dst + base + lane + stride * j, v);
    }
}

struct FinalArgs {
    const uint4* src;
    uint4* dst;
    uint64_t src_pitch, dst_pitch;  // per column, in elements
    uint32_t log_n, log_s;          // S = n_1
    uint32_t passes;                // p in {1,2,3}
    uint32_t log_top;               // log2(n_p), the most significant storage digit (p >= 2)
    uint32_t mode;                  // 0 plain, 1 LDE from tmp (8 cosets/tile), 2 LDE direct from coefficients (p = 1)
    uint32_t lanes_log;             // 0 (one run per tile) or 3
    uint32_t log_L;
    const uint4* roots;
    const uint4* off_tab;
    NttScale scale;
};

__global__ void __launch_bounds__(kThreads) ntt_final_pass(FinalArgs a) {
    extern __shared__ uint4 tile[];
    const uint32_t S = 1u << a.log_s;
    const uint32_t lanes = 1u << a.lanes_log;
    const uint32_t pitch = lanes > 1 ? kPad : 1;
    const uint32_t log_H = a.log_n - a.log_s;  // runs per column
    const uint32_t col = blockIdx.y;
    // which runs does this tile own?
    uint64_t run0, run_step;  // run index of lane r = run0 + r * run_step
    uint64_t out_base;        // output index (before the j_1 term) of lane 0
    if (a.mode == 0) {
        if (a.passes == 1) {
            run0 = 0, run_step = 0, out_base = 0;
        } else {
            // hi' = rest + (H / n_p) * j_p ; tile owns j_p = jp0 .. jp0+7 for one `rest`
            const uint32_t log_rest = log_H - a.log_top;
            const uint32_t tiles_per_rest = (1u << a.log_top) / kLanes;
            const uint64_t rest = blockIdx.x / tiles_per_rest;
            const uint32_t jp0 = (blockIdx.x % tiles_per_rest) * kLanes;
            run0 = rest + ((uint64_t)jp0 << log_rest);
            run_step = 1ull << log_rest;
            out_base = (rest << a.log_top) + jp0;  // j_p + n_p * rest
        }
    } else {
        // LDE: one run (hi') per tile, the 8 lanes are the cosets
        const uint64_t hp = blockIdx.x;
        run0 = hp, run_step = 0;
        if (a.passes <= 1) {
            out_base = 0;
        } else {
            const uint32_t log_rest = log_H - a.log_top;
            const uint64_t rest = hp & ((1ull << log_rest) - 1), jp = hp >> log_rest;
            out_base = (rest << a.log_top) + jp;
        }
    }
    const uint32_t total = S << a.lanes_log;
    // load: lanes are separate runs; consecutive threads read consecutive elements of one run
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        uint32_t m = e & (S - 1), lane = e >> a.log_s;
        fe v;
        if (a.mode == 0) {
            const uint4* src = a.src + (uint64_t)col * a.src_pitch;
            v = fe_load(src + ((run0 + lane * run_step) << a.log_s) + m);
        } else if (a.mode == 1) {
            const uint4* src = a.src + ((uint64_t)col * 8 + lane) * a.src_pitch;
            v = fe_load(src + (run0 << a.log_s) + m);
        } else {
            const uint4* src = a.src + (uint64_t)col * a.src_pitch;
            v = fe_load(src + m);
            if (lane != 0) v = fe_mul(v, fe_root_pow(a.roots, a.log_L, (uint64_t)lane * m));
        }
        fe_store(tile + m * pitch + lane, v);
    }
    __syncthreads();
    tile_dif(tile, a.log_s, a.lanes_log, pitch, a.roots);
    uint4* dst = a.dst + (uint64_t)col * a.dst_pitch;
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        uint32_t lane = e & (lanes - 1), pos = e >> a.lanes_log;
        uint32_t j1 = bitrev32(pos, a.log_s);
        fe v = fe_load(tile + pos * pitch + lane);
        uint64_t out;
        if (a.mode == 0) {
            out = ((uint64_t)j1 << log_H) + out_base + lane;
            if (a.scale.enabled) {
                uint32_t ci = (uint32_t)(out >> a.scale.chunk_shift);
                fe c = fe_make(a.scale.cvec[ci][0], a.scale.cvec[ci][1]);
                if (a.scale.use_offset) c = fe_mul(c, fe_tab_pow(a.off_tab, (uint32_t)out));
                v = fe_mul(v, c);
            }
        } else {
            out = ((((uint64_t)j1 << log_H) + out_base) << 3) + lane;
        }
        fe_store(dst + out, v);
    }
}

size_t tile_bytes(uint32_t log_s, uint32_t lanes, bool padded) {
    return ((size_t)1 << log_s) * (lanes > 1 ? (padded ? kPad : kLanes) : 1) * sizeof(uint4);
}

struct Plan {
    uint32_t passes;
    uint32_t log_d[3];  // log n_1, n_2, n_3
};

Plan make_plan(uint32_t log_n, int max_tile_log) {
    Plan p{};
    const uint32_t T = (uint32_t)max_tile_log;
    if (log_n <= T) {
        p.passes = 1, p.log_d[0] = log_n;
    } else if (log_n <= 2 * T) {
        p.passes = 2, p.log_d[0] = (log_n + 1) / 2, p.log_d[1] = log_n / 2;
    } else {
        p.passes = 3;
        p.log_d[0] = (log_n + 2) / 3;
        uint32_t rest = log_n - p.log_d[0];
        p.log_d[1] = (rest + 1) / 2, p.log_d[2] = rest / 2;
    }
    return p;
}

bool g_attr_set = false;
void ensure_smem_attr() {
    if (g_attr_set) return;
    EZK_CUDA(cudaFuncSetAttribute(ntt_strided_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    EZK_CUDA(cudaFuncSetAttribute(ntt_final_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    g_attr_set = true;
}

// strided passes p..2 over `ncols` arrays of n elements: the first pass reads `first_src` and writes `buf`,
// later passes run in place in `buf`.  With `coset` the first pass reads coefficient column y/8 and applies the
// coset factor w_L^(c*m), c = y%8 (LDE).
int run_strided(const NttTables& t, cudaStream_t s, const Plan& pl, uint32_t log_n, const uint4* roots,
                const uint4* first_src, uint64_t first_src_pitch, bool coset, uint4* buf, uint64_t buf_pitch,
                uint32_t ncols, uint32_t log_L) {
    int launches = 0;
    uint32_t log_stride = log_n;
    for (int i = (int)pl.passes - 1; i >= 1; i--) {
        log_stride -= pl.log_d[i];
        StridedArgs a{};
        bool first = (i == (int)pl.passes - 1) && first_src != nullptr;
        a.src = first ? first_src : buf;
        a.src_pitch = first ? first_src_pitch : buf_pitch;
        a.dst = buf, a.dst_pitch = buf_pitch;
        a.log_stride = log_stride, a.log_s = pl.log_d[i];
        a.coset_first = (first && coset) ? 1 : 0;
        a.log_L = log_L;
        a.roots = roots;
        dim3 grid((unsigned)(((uint64_t)1 << (log_n - pl.log_d[i])) / kLanes), ncols);
        {
            // compulsory traffic: every element of every column read once and written once (the 8 coset copies of
            // an LDE first pass share one read of the coefficient column)
            const uint64_t elems = (uint64_t)ncols << log_n;
            LaunchScope ls(s, K_NTT_STRIDED, (a.coset_first ? elems / 8 + elems : 2 * elems) * 16);
            ntt_strided_pass<<<grid, kThreads, tile_bytes(pl.log_d[i], kLanes, false), s>>>(a);
        }
        EZK_CUDA(cudaGetLastError());
        launches++;
    }
    (void)t;
    return launches;
}

}  // namespace

void ntt_tables_init(NttTables& t) {
    auto build = [](Fp base) {
        std::vector<Fp> h(2 * EZK_TAB_SIZE);
        Fp acc(1);
        for (uint32_t k = 0; k < EZK_TAB_SIZE; k++) {
            h[k] = acc;
            acc = acc * base;
        }
        Fp big = acc;  // base^(2^14)
        acc = Fp(1);
        for (uint32_t k = 0; k < EZK_TAB_SIZE; k++) {
            h[EZK_TAB_SIZE + k] = acc;
            acc = acc * big;
        }
        uint4* d = nullptr;
        EZK_CUDA(cudaMalloc(&d, h.size() * sizeof(uint4)));
        EZK_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        return d;
    };
    Fp w = root_of_unity(EZK_ROOT_LOG), o = Fp::from_u64(kDomainOffset);
    t.root_fwd = build(w);
    t.root_inv = build(inverse(w));
    t.off_fwd = build(o);
    t.off_inv = build(inverse(o));
    const char* env = getenv("EZK_NTT_TILE_LOG");
    if (env) {
        int v = atoi(env);
        if (v >= 9 && v <= 10) t.max_tile_log = v;
    }
}

void ntt_tables_free(NttTables& t) {
    cudaFree(t.root_fwd), cudaFree(t.root_inv), cudaFree(t.off_fwd), cudaFree(t.off_inv);
    t = NttTables{};
}

int ntt_columns(const NttTables& t, cudaStream_t s, const uint4* src, uint64_t src_pitch, uint4* dst, uint64_t dst_pitch,
                uint4* work, uint32_t ncols, uint32_t log_n, bool inverse, const NttScale* scale) {
    ensure_smem_attr();
    const uint4* roots = inverse ? t.root_inv : t.root_fwd;
    Plan pl = make_plan(log_n, t.max_tile_log);
    const uint64_t n = 1ull << log_n;
    int launches = 0;
    FinalArgs a{};
    if (pl.passes >= 2) {
        launches = run_strided(t, s, pl, log_n, roots, src, src_pitch, false, work, n, ncols, 0);
        a.src = work, a.src_pitch = n;
    } else {
        a.src = src, a.src_pitch = src_pitch;
    }
    a.dst = dst, a.dst_pitch = dst_pitch;
    a.log_n = log_n, a.log_s = pl.log_d[0], a.passes = pl.passes;
    a.log_top = pl.passes >= 2 ? pl.log_d[pl.passes - 1] : 0;
    a.mode = 0;
    a.lanes_log = pl.passes >= 2 ? 3 : 0;
    a.roots = roots;
    if (scale) {
        a.scale = *scale;
        a.scale.enabled = 1;
        a.off_tab = scale->use_offset == 2 ? t.off_inv : t.off_fwd;
    }
    uint64_t runs = 1ull << (log_n - pl.log_d[0]);
    dim3 grid((unsigned)(pl.passes >= 2 ? runs / kLanes : 1), ncols);
    {
        LaunchScope ls(s, K_NTT_FINAL, ((uint64_t)ncols << log_n) * 32);
        ntt_final_pass<<<grid, kThreads, tile_bytes(pl.log_d[0], 1u << a.lanes_log, true), s>>>(a);
    }
    EZK_CUDA(cudaGetLastError());
    return launches + 1;
}

int lde_columns(const NttTables& t, cudaStream_t s, const uint4* coeff, uint64_t coeff_pitch, uint4* lde,
                uint64_t lde_pitch, uint4* tmp, uint32_t ncols, uint32_t log_n) {
    ensure_smem_attr();
    const uint4* roots = t.root_fwd;
    Plan pl = make_plan(log_n, t.max_tile_log);
    const uint64_t n = 1ull << log_n;
    const uint32_t log_L = log_n + 3;
    int launches = 0;
    FinalArgs a{};
    a.dst = lde, a.dst_pitch = lde_pitch;
    a.log_n = log_n, a.log_s = pl.log_d[0], a.passes = pl.passes;
    a.log_top = pl.passes >= 2 ? pl.log_d[pl.passes - 1] : 0;
    a.lanes_log = 3;
    a.log_L = log_L;
    a.roots = roots;
    if (pl.passes == 1) {
        a.mode = 2;
        a.src = coeff, a.src_pitch = coeff_pitch;
    } else {
        launches += run_strided(t, s, pl, log_n, roots, coeff, coeff_pitch, true, tmp, n, ncols * 8, log_L);
        a.mode = 1;
        a.src = tmp, a.src_pitch = n;
    }
    dim3 grid((unsigned)(1ull << (log_n - pl.log_d[0])), ncols);
    {
        const uint64_t elems = (uint64_t)ncols << log_n;
        LaunchScope ls(s, K_NTT_FINAL, (pl.passes == 1 ? elems + 8 * elems : 16 * elems) * 16);
        ntt_final_pass<<<grid, kThreads, tile_bytes(pl.log_d[0], 8, true), s>>>(a);
    }
    EZK_CUDA(cudaGetLastError());
    return launches + 1;
}

}  // namespace ezk
