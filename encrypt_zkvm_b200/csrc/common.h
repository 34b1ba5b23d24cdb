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

inline unsigned ilog2_u64(uint64_t n) {
    unsigned k = 0;
    while ((1ull << k) < n) k++;
    return k;
}

}  // namespace ezk
