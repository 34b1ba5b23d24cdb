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

inline unsigned ilog2_u64(uint64_t n) {
    unsigned k = 0;
    while ((1ull << k) < n) k++;
    return k;
}

}  // namespace ezk
