// BLAKE3 row hashing + Merkle tree kernels (K3, K4).  See merkle.cuh.
#include "merkle.cuh"
#include "../common.h"
#include "../hash/blake3.cuh"
#include <algorithm>

namespace ezk {

using namespace dev;

namespace {

constexpr int kThreads = 256;

// One thread per row; the table is column-major so a warp reads 32 consecutive 16-byte cells per column.
// WIDTH > 0: the row width is a compile-time constant (28 trace columns, 7 composition columns, 8 FRI values), so
// the block loop is unrolled and the block lengths and flags are immediates; WIDTH = 0: any width.
// Where the digest of packed row t goes.  LocalLeaves: leaves[t] of this GPU.  PeerSlots (multi-GPU, fused exchange):
// row t of rank `me` belongs to the subtree of rank q = t / chunk and is stored straight into that rank's receive
// area over NVLink (peer pointers from Comm::map_peers), at [me][t mod chunk] - the layout an all-to-all would have
// produced, so consecutive threads write consecutive 32-byte digests (scattering them to their final, interleaved leaf
// slots made 32-byte writes at a stride of world * 32 bytes: 1.3 -> 2.0 ms of hash kernels per proof on 8 GPUs).
struct LocalLeaves {
    uint4* leaves;
    __device__ __forceinline__ uint4* slot(uint64_t t) const { return leaves + 2 * t; }
};
struct PeerSlots {
    uint4* recv[8];  // receive area of every rank: world chunks of `chunk` digests, one per sender
    uint32_t chunk_log, me;
    __device__ __forceinline__ uint4* slot(uint64_t t) const {
        return recv[t >> chunk_log] + 2 * (((uint64_t)me << chunk_log) + (t & ((1ull << chunk_log) - 1)));
    }
};

template <int WIDTH, class Sink>
__global__ void __launch_bounds__(kThreads) hash_rows_kernel(const uint4* __restrict__ table, uint64_t pitch,
                                                            uint32_t width_rt, uint64_t rows, Sink sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    const uint32_t width = WIDTH > 0 ? (uint32_t)WIDTH : width_rt;
    uint32_t cv[8];
    b3_init(cv);
    const uint32_t nblocks = (width + 3) / 4;
    const uint4* cell = table + t;  // multi-GPU: tables hold this rank's rows in packed order
#pragma unroll
    for (uint32_t b = 0; b < (WIDTH > 0 ? (uint32_t)(WIDTH + 3) / 4 : nblocks); b++) {
        uint32_t m[16];
        const uint32_t cells = min(4u, width - 4 * b);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (k < cells) v = __ldg(cell + (uint64_t)(4 * b + k) * pitch);
            m[4 * k] = v.x, m[4 * k + 1] = v.y, m[4 * k + 2] = v.z, m[4 * k + 3] = v.w;
        }
        const uint32_t flags = (b == 0 ? B3_CHUNK_START : 0u) | (b == nblocks - 1 ? (B3_CHUNK_END | B3_ROOT) : 0u);
        b3_compress(cv, m, cells * 16, flags);
    }
    uint4* out = sink.slot(t);
    out[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    out[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

// dst[global_row_q(t) * UNITS + u] = src[(q * per_rank + t) * UNITS + u]: puts the all-gathered per-rank blocks
// (packed row order) back into natural row order.  UNITS = 16-byte words per item (1 element, 2 digest).
template <int UNITS>
__global__ void unpack_rows_kernel(const uint4* __restrict__ src, uint64_t per_rank, uint32_t world_log, uint4* __restrict__ dst) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (per_rank << world_log)) return;
    RowShard sh;
    sh.rank = (uint32_t)(g / per_rank), sh.world_log = world_log;
    const uint64_t i = sh.global_row(g % per_rank);
#pragma unroll
    for (int u = 0; u < UNITS; u++) dst[i * UNITS + u] = src[g * UNITS + u];
}

// `depth` (<= 9) levels per launch: CTA b merges the 512 digests under parents [256 b, 256 b + 256) of the level with
// `level` parents, keeps the results in shared memory and goes on with the 128, 64, ... parents above them; every level
// is also written to the node array.  level must be a multiple of 256.
__global__ void __launch_bounds__(kThreads) merkle_subtree_kernel(uint4* __restrict__ nodes, uint64_t level, int depth) {
    __shared__ uint4 sm[2][2 * kThreads];
    const uint32_t t = threadIdx.x;
    uint4 out[2];
    {
        const uint64_t k = level + (uint64_t)blockIdx.x * kThreads + t;
        b3_merge(nodes + 4 * k, out);
        nodes[2 * k] = out[0], nodes[2 * k + 1] = out[1];
        sm[0][2 * t] = out[0], sm[0][2 * t + 1] = out[1];
    }
    __syncthreads();
    uint32_t width = kThreads;
    int cur = 0;
    for (int d = 1; d < depth; d++) {
        level >>= 1, width >>= 1;
        if (t < width) {
            b3_merge(&sm[cur][4 * t], out);
            const uint64_t k = level + (uint64_t)blockIdx.x * width + t;
            nodes[2 * k] = out[0], nodes[2 * k + 1] = out[1];
            sm[cur ^ 1][2 * t] = out[0], sm[cur ^ 1][2 * t + 1] = out[1];
        }
        __syncthreads();
        cur ^= 1;
    }
}

// finishes the tree from `level` (<= 1024 parents) down to the root inside one CTA
__global__ void __launch_bounds__(1024) merkle_top_kernel(uint4* __restrict__ nodes, uint32_t level) {
    for (uint32_t lvl = level; lvl >= 1; lvl >>= 1) {
        if (threadIdx.x < lvl) {
            uint64_t k = lvl + threadIdx.x;
            uint4 out[2];
            b3_merge(nodes + 4 * k, out);
            nodes[2 * k] = out[0];
            nodes[2 * k + 1] = out[1];
        }
        __syncthreads();
    }
}

__global__ void gather_rows_kernel(const uint4* __restrict__ table, uint64_t pitch, uint32_t width,
                                   const uint64_t* __restrict__ idx, uint32_t nq, uint4* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * width) return;
    uint32_t q = t / width, c = t % width;
    out[t] = table[(uint64_t)c * pitch + idx[q]];
}

__global__ void gather_digests_kernel(const uint4* __restrict__ nodes, const uint64_t* __restrict__ idx, uint32_t nq,
                                      uint4* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    out[2 * t] = nodes[2 * idx[t]];
    out[2 * t + 1] = nodes[2 * idx[t] + 1];
}

}  // namespace

int merkle_hash_rows(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t rows, uint4* nodes) {
    return hash_rows_sharded(s, table, pitch, width, rows, RowShard(), nodes + 2 * rows);
}

template <class Sink>
static void launch_hash_rows(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t rows, const Sink& sink) {
    unsigned blocks = (unsigned)((rows + kThreads - 1) / kThreads);
    {
        LaunchScope ls(s, K_HASH_ROWS, rows * ((uint64_t)width * 16 + 32));
        if (width == 28)
            hash_rows_kernel<28><<<blocks, kThreads, 0, s>>>(table, pitch, width, rows, sink);
        else if (width == 7)
            hash_rows_kernel<7><<<blocks, kThreads, 0, s>>>(table, pitch, width, rows, sink);
        else if (width == 8)
            hash_rows_kernel<8><<<blocks, kThreads, 0, s>>>(table, pitch, width, rows, sink);
        else
            hash_rows_kernel<0><<<blocks, kThreads, 0, s>>>(table, pitch, width, rows, sink);
    }
    EZK_CUDA(cudaGetLastError());
}

int hash_rows_sharded(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t local_rows, RowShard,
                      uint4* digests) {
    launch_hash_rows(s, table, pitch, width, local_rows, LocalLeaves{digests});
    return 1;
}

int hash_rows_to_peers(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t local_rows, RowShard sh,
                       uint4* const peer_recv[8]) {
    PeerSlots sink;
    for (int q = 0; q < 8; q++) sink.recv[q] = peer_recv[q];
    sink.me = sh.rank;
    sink.chunk_log = ilog2_u64(local_rows) - sh.world_log;  // rows per destination = local_rows / world
    launch_hash_rows(s, table, pitch, width, local_rows, sink);
    return 1;
}

int unpack_rows(cudaStream_t s, const uint4* gathered, uint64_t per_rank, uint32_t world_log, uint32_t units, uint4* dst) {
    const uint64_t total = per_rank << world_log;
    unsigned blocks = (unsigned)((total + 255) / 256);
    {
        LaunchScope ls(s, K_GATHER, total * units * 32);
        if (units == 2)
            unpack_rows_kernel<2><<<blocks, 256, 0, s>>>(gathered, per_rank, world_log, dst);
        else
            unpack_rows_kernel<1><<<blocks, 256, 0, s>>>(gathered, per_rank, world_log, dst);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int merkle_build(cudaStream_t s, uint4* nodes, uint64_t num_leaves) {
    int launches = 0;
    uint64_t level = num_leaves / 2;
    while (level > 1024) {
        // up to 9 levels per launch (a CTA carries its 256 parents up through shared memory); the last 1024 parents
        // and everything above them belong to the single-CTA top kernel
        const int depth = (int)std::min<unsigned>(9, ilog2_u64(level) - 10);
        {
            LaunchScope ls(s, K_MERKLE_LEVEL, level * 96 * 2);
            merkle_subtree_kernel<<<(unsigned)(level / kThreads), kThreads, 0, s>>>(nodes, level, depth);
        }
        EZK_CUDA(cudaGetLastError());
        launches++;
        level >>= depth;
    }
    if (level >= 1) {
        {
            LaunchScope ls(s, K_MERKLE_TOP, level * 2 * 96);
            merkle_top_kernel<<<1, 1024, 0, s>>>(nodes, (uint32_t)level);
        }
        EZK_CUDA(cudaGetLastError());
        launches++;
    }
    return launches;
}

int gather_rows(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, const uint64_t* idx, uint32_t nq,
                uint4* out) {
    unsigned total = nq * width;
    {
        LaunchScope ls(s, K_GATHER, (uint64_t)total * 32);
        gather_rows_kernel<<<(total + 127) / 128, 128, 0, s>>>(table, pitch, width, idx, nq, out);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int gather_digests(cudaStream_t s, const uint4* nodes, const uint64_t* idx, uint32_t nq, uint4* out) {
    {
        LaunchScope ls(s, K_GATHER, (uint64_t)nq * 64);
        gather_digests_kernel<<<(nq + 127) / 128, 128, 0, s>>>(nodes, idx, nq, out);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace ezk
