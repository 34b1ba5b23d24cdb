// Shared host-side helpers for the CUDA translation units: error propagation and launch counting.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>

namespace ezk {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define EZK_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::ezk::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                                   __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
    } while (0)

// number of kernels launched by this library since load (reported by bench.py as gpu_launches)
void count_launch(uint64_t k = 1);
uint64_t launch_count();

// Per-kernel device timing (CUDA events on the launching stream) and algorithmic byte accounting, switched on by
// ezk_profile_enable(); used by bench.py for the roofline line.  Off by default: no events are recorded.
enum KernelId {
    K_NTT_STRIDED = 0,
    K_NTT_FINAL,
    K_HASH_ROWS,
    K_MERKLE_LEVEL,
    K_MERKLE_TOP,
    K_GATHER,
    K_PAIR_INVERSE,
    K_CONSTRAINTS,
    K_FRAMES,
    K_EVAL_POLYS,
    K_DEEP_COMBINE,
    K_DEEP_POINTWISE,
    K_ALL_ZERO,
    K_FRI_FOLD,
    K_FRI_REMAINDER,
    K_COUNT
};
const char* kernel_name(int id);
void profile_enable(bool on);
void profile_reset();
// totals since the last reset; synchronises the device to read the events
void profile_read(int id, uint64_t* launches, double* ms, uint64_t* algo_bytes);

// Brackets one kernel launch: counts it and, when profiling, times it and books its algorithmic bytes.
struct LaunchScope {
    LaunchScope(cudaStream_t s, KernelId id, uint64_t algo_bytes);
    ~LaunchScope();
    cudaStream_t stream;
    int slot;
};

// Multi-GPU row ownership: rank r of 2^world_log ranks owns the LDE rows i with i mod world = r, i.e. whole
// cosets, so the (row i, row i + 8) frames of the constraint evaluation and the 8 points of a FRI layer-0 row stay on
// one GPU.  A rank keeps its rows in ascending order ("packed": row i = world * t + r at index t < L / world), tables
// included; row kernels run over t, take the domain point from global_row(t), and find the row 8 steps ahead at
// t + 8 / world.  With one GPU global_row is the identity.
struct RowShard {
    uint32_t rank = 0, world_log = 0;
    __host__ __device__ uint64_t global_row(uint64_t t) const {
        const uint32_t cn_log = 3 - world_log;  // log2(cosets per rank)
        return ((t >> cn_log) << 3) + rank + ((t & ((1u << cn_log) - 1)) << world_log);
    }
    __host__ __device__ uint32_t owner(uint64_t i) const { return (uint32_t)(i & ((1u << world_log) - 1)); }
    __host__ __device__ uint32_t frame_step() const { return 8u >> world_log; }  // packed distance of LDE rows i and i + 8
};

inline unsigned ilog2_u64(uint64_t n) {
    unsigned k = 0;
    while ((1ull << k) < n) k++;
    return k;
}

}  // namespace ezk
