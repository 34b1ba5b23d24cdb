// See expand.cuh.
#include "expand.cuh"
#include "../common.h"

namespace ezk {

namespace {

constexpr int kChunk = 1024;

// depth change of one operation (vm/src/processor/stack.rs:48-70 through shift_left / shift_right); 127 = unknown opcode
__device__ __forceinline__ int depth_delta(uint32_t code, int lw) {
    switch (code) {
        case 0b00000: return 0;        // noop
        case 0b10000: return 1;        // push
        case 0b10001: return 1;        // read
        case 0b10010: return lw;       // read2
        case 0b01000: return -1;       // add
        case 0b01001: return -1;       // mul
        case 0b01010: return -1;       // sadd: scalar + ciphertext -> ciphertext
        case 0b01100: return -1;       // smul
        case 0b01011: return -lw;      // add2
    }
    return 127;
}

// sums[b] = sum of the depth changes of operations [1024 b, 1024 b + 1024)
__global__ void __launch_bounds__(256) chunk_sums_kernel(const uint8_t* __restrict__ codes, uint64_t count, int lw,
                                                        uint32_t* __restrict__ sums, uint32_t* __restrict__ flag) {
    __shared__ int red[256];
    const uint64_t base = (uint64_t)blockIdx.x * kChunk;
    int acc = 0;
    bool bad = false;
    for (int k = threadIdx.x; k < kChunk; k += 256) {
        const uint64_t i = base + k;
        if (i < count) {
            const int d = depth_delta(codes[i], lw);
            bad |= d == 127;
            acc += d == 127 ? 0 : d;
        }
    }
    if (bad) atomicOr(flag, 4u);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int h = 128; h >= 1; h >>= 1) {
        if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = (uint32_t)red[0];
}

// exclusive scan of the chunk sums in place (one CTA; at most 2^14 + 1 chunks)
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ sums, uint32_t chunks) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (chunks + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per, hi = min(chunks, lo + per);
    uint32_t acc = 0;
    for (uint32_t k = lo; k < hi; k++) acc += sums[k];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele over the per-thread totals
        const uint32_t v = (int)threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (uint32_t k = lo; k < hi; k++) {
        const uint32_t v = sums[k];
        sums[k] = run;
        run += v;
    }
}

__device__ __forceinline__ uint4 small(uint32_t v) { return make_uint4(v, 0, 0, 0); }

// one CTA per chunk of 1024 rows: local scan of the depth changes, then the eight columns
__global__ void __launch_bounds__(256) write_columns_kernel(const uint8_t* __restrict__ codes, uint64_t count, uint64_t n, int lw,
                                                           const uint32_t* __restrict__ offsets, const uint4* __restrict__ last_row,
                                                           uint4* __restrict__ trace) {
    __shared__ int pre[kChunk + 1];
    __shared__ int tot[256];
    const uint64_t base = (uint64_t)blockIdx.x * kChunk;
    // each thread owns 4 consecutive rows of the chunk
    int d[4], acc = 0;
    uint32_t code[4];
    for (int k = 0; k < 4; k++) {
        const uint64_t i = base + 4 * threadIdx.x + k;
        code[k] = i < count ? codes[i] : 0;
        const int dd = i < count ? depth_delta(code[k], lw) : 0;
        d[k] = dd == 127 ? 0 : dd;
        acc += d[k];
    }
    tot[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int v = (int)threadIdx.x >= off ? tot[threadIdx.x - off] : 0;
        __syncthreads();
        tot[threadIdx.x] += v;
        __syncthreads();
    }
    int run = (int)offsets[blockIdx.x] + (threadIdx.x ? tot[threadIdx.x - 1] : 0);  // depth BEFORE this thread's first op
    for (int k = 0; k < 4; k++) {
        pre[4 * threadIdx.x + k] = run;  // depth at row i = depth after operations 0 .. i-1
        run += d[k];
    }
    __syncthreads();
    for (int k = 0; k < 4; k++) {
        const uint64_t i = base + 4 * threadIdx.x + k;
        if (i >= n) break;
        if (i == n - 1) {  // the random row (mod.rs:86-92): caller-supplied for every column
            for (int c = 0; c < 7; c++) trace[(uint64_t)c * n + i] = last_row[c];
            trace[11ull * n + i] = last_row[11];
            continue;
        }
        const bool live = i < count;
        trace[i] = make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0, 0);
        for (int b = 0; b < 5; b++) trace[(uint64_t)(1 + b) * n + i] = small(live ? (code[k] >> b) & 1u : 0u);
        trace[6ull * n + i] = small(live ? 1u : 0u);
        trace[11ull * n + i] = small((uint32_t)pre[4 * threadIdx.x + k]);  // rows past the program repeat the last depth
    }
}

}  // namespace

int expand_op_columns(cudaStream_t s, const uint8_t* d_codes, uint64_t count, uint64_t n, uint32_t lwe_size,
                      const uint4* d_last_row, uint32_t* d_scan, uint4* d_trace, uint32_t* d_flag) {
    const uint32_t chunks = (uint32_t)((n + kChunk - 1) / kChunk);
    {
        LaunchScope ls(s, K_GATHER, count);
        chunk_sums_kernel<<<chunks, 256, 0, s>>>(d_codes, count, (int)lwe_size, d_scan, d_flag);
    }
    EZK_CUDA(cudaGetLastError());
    {
        LaunchScope ls(s, K_GATHER, (uint64_t)chunks * 8);
        scan_sums_kernel<<<1, 1024, 0, s>>>(d_scan, chunks);
    }
    EZK_CUDA(cudaGetLastError());
    {
        LaunchScope ls(s, K_GATHER, count + n * 8 * 16);
        write_columns_kernel<<<chunks, 256, 0, s>>>(d_codes, count, n, (int)lwe_size, d_scan, d_last_row, d_trace);
    }
    EZK_CUDA(cudaGetLastError());
    return 3;
}

}  // namespace ezk
