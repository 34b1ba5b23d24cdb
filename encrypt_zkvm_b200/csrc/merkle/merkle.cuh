// BLAKE3 row hashing + Merkle tree construction on the GPU (kernels K3, K4 of SURVEY 8a').
// Replaces winter-crypto's `Blake3_256::hash_elements` over LDE rows and `MerkleTree::new`, as used by
// DefaultTraceLde::new (prover/src/lib.rs:55-62), the constraint commitment and FRI layers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../common.h"

namespace ezk {

// Node array layout (same as winter-crypto): `nodes` holds 2*num_leaves digests of 32 bytes;
// nodes[num_leaves + i] = leaf i, nodes[k] = blake3(nodes[2k] || nodes[2k+1]), nodes[1] = root.

// leaf i = blake3(row i as little-endian bytes), row i = (table[c * pitch + i])_{c < width}
int merkle_hash_rows(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t rows, uint4* nodes);
// multi-GPU: digests[t] = blake3(row t of the table) for the local_rows rows this rank holds (packed order: table row t
// is LDE row global_row(t))
int hash_rows_sharded(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t local_rows, RowShard sh,
                      uint4* digests);
// multi-GPU, exchange fused into the hash: the digest of this rank's packed row t (LDE row world * t + rank) is stored
// straight into the receive area of the rank whose subtree it belongs to (peer_recv[q] = rank q's receive area,
// reachable over NVLink), at [rank][t mod chunk], chunk = local_rows / world - what an all-to-all would deliver.
// local_rows must be a power of two >= world.  Callers put a group-wide barrier between this launch and unpack_rows.
int hash_rows_to_peers(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, uint64_t local_rows, RowShard sh,
                       uint4* const peer_recv[8]);
// dst[global_row_q(t)] = gathered[q][t] for the 2^world_log all-gathered blocks of per_rank items (units x 16 bytes each)
int unpack_rows(cudaStream_t s, const uint4* gathered, uint64_t per_rank, uint32_t world_log, uint32_t units, uint4* dst);
// builds all internal nodes from the leaves already stored in nodes[num_leaves..2*num_leaves)
int merkle_build(cudaStream_t s, uint4* nodes, uint64_t num_leaves);

// out[q * width + c] = table[c * pitch + idx[q]]  (row gather for query openings)
int gather_rows(cudaStream_t s, const uint4* table, uint64_t pitch, uint32_t width, const uint64_t* idx, uint32_t nq,
                uint4* out);
// out[q] = nodes[idx[q]]  (digest gather for authentication paths)
int gather_digests(cudaStream_t s, const uint4* nodes, const uint64_t* idx, uint32_t nq, uint4* out);

}  // namespace ezk
