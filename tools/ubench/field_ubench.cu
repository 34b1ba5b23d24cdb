// Micro-benchmarks of the f128 primitives on one GPU: issue cost (cycles per warp-instruction-group per SM
// sub-partition) of the modular product, sum and difference, with 1..8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o field_ubench field_ubench.cu && ./field_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../encrypt_zkvm_b200/csrc/field/f128.cuh"
using namespace ezk::dev;

template <int OP, int CHAINS>
__global__ void __launch_bounds__(256) kern(uint4* out, int iters, uint4 seed) {
    fe x[CHAINS], w = fe_from(seed);
    fe_pre wp;
    wp.w[0] = w, wp.w[1] = fe_make(seed.y * 3u + 1u, seed.z), wp.w[2] = fe_make(seed.x + 7u, seed.w), wp.w[3] = fe_make(seed.z, seed.x);
    for (int c = 0; c < CHAINS; c++) x[c] = fe_make(threadIdx.x * 977u + c * 131u + 5u, blockIdx.x + 3u + c);
    uint32_t rare = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == 0) x[c] = fe_mul_flag(x[c], w, rare);
            if (OP == 1) x[c] = fe_add_flag(x[c], w, rare);
            if (OP == 2) x[c] = fe_sub(x[c], w);
            if (OP == 3) x[c] = fe_mul(x[c], w);
            if (OP == 5) x[c] = fe_mul_pre_flag(x[c], wp, rare);  // product with a table constant in precomputed form
            if (OP == 6) {  // NTT butterfly: add, sub, precomputed-form twiddle product
                fe s = fe_add_flag(x[c], w, rare), d = fe_sub(x[c], w);
                x[c] = fe_add_flag(s, fe_mul_pre_flag(d, wp, rare), rare);
            }
            if (OP == 4) {  // butterfly-like mix: 1 mul + 1 add + 1 sub
                fe s = fe_add_flag(x[c], w, rare), d = fe_sub(x[c], w);
                x[c] = fe_add_flag(s, fe_mul_flag(d, w, rare), rare);
            }
        }
    }
    fe acc = fe_zero();
    for (int c = 0; c < CHAINS; c++) acc = fe_add(acc, x[c]);
    if (rare == 0x12345u) acc.a0 ^= 1;
    out[blockIdx.x * blockDim.x + threadIdx.x] = fe_to(acc);
}

template <int OP, int CHAINS>
void run(const char* name, int blocks_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * blocks_per_sm, iters = 2000;
    uint4* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 16);
    uint4 seed = make_uint4(0x9E3779B9u, 0x7F4A7C15u, 0xF39CC060u, 0x5CEDC834u);
    kern<OP, CHAINS><<<blocks, 256>>>(out, 10, seed);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<OP, CHAINS><<<blocks, 256>>>(out, iters, seed);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * 256 * iters * CHAINS;  // thread-level ops
    const double warp_ops_per_smsp = ops / 32 / (sms * 4);
    printf("%-10s chains=%d blocks/SM=%d: %8.3f ms  %7.1f Gop/s  %6.1f ns*GHz-free cycles@1.9GHz per warp-op per SMSP = %.1f\n", name, CHAINS,
           blocks_per_sm, ms, ops / ms / 1e6, 0.0, ms * 1e-3 * 1.9e9 / warp_ops_per_smsp);
    cudaFree(out);
}

int main() {
    run<0, 1>("mul_flag", 4);
    run<0, 2>("mul_flag", 4);
    run<0, 4>("mul_flag", 4);
    run<0, 8>("mul_flag", 4);
    run<0, 8>("mul_flag", 2);
    run<0, 8>("mul_flag", 3);
    run<3, 8>("mul_exact", 4);
    run<1, 8>("add_flag", 4);
    run<2, 8>("sub", 4);
    run<5, 8>("mul_pre", 4);
    run<5, 4>("mul_pre", 4);
    run<6, 8>("bfly_pre", 4);
    run<4, 8>("bfly", 4);
    run<4, 8>("bfly", 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
