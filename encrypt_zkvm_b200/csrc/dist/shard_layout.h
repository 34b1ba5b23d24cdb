// Index arithmetic of the sharded proof, host side.  Shared by the prover's opening planner and by the
// host-only self-test ezk_selftest_shard_layout, so the routing rules are exercised on a CPU-only box.
//
//   rows      rank r of G owns the rows i = G t + r and keeps them at index t ("packed", ascending).
//   leaves    the Merkle tree of a commitment with `leaves` leaves is cut at depth log2 G: rank q builds the subtree
//             over leaves [q ll, (q + 1) ll), ll = leaves / G.  Rank r's packed digest t belongs to leaf G t + r, i.e. to
//             subtree (G t + r) / ll = t / (ll / G): its digests are already grouped by destination, chunk q =
//             t in [q ll / G, (q + 1) ll / G)  ->  one all-to-all with equal chunks.  Rank q receives from rank r' the
//             chunk whose entry t' is leaf q ll + G t' + r' : local leaf G t' + r'.
//   nodes     heap order, node 1 = root, leaves at [leaves, 2 leaves).  Nodes below the cut live in a subtree (local heap
//             index, local root = 1); nodes [1, 2G) - the G subtree roots and the levels above - live on the host.
#pragma once
#include <cstdint>

namespace ezk {

struct NodeHome {
    int owner;       // rank that holds the node, or -1: host-side top array (index = global heap index < 2 G)
    uint64_t index;  // local heap index inside the owner's subtree
};

inline unsigned floor_log2_u64(uint64_t v) {
    unsigned k = 0;
    while (v >>= 1) k++;
    return k;
}

// home of global node k (k >= 1) of a tree split over G = 2^glog subtrees
inline NodeHome shard_node_home(uint32_t G, uint32_t glog, uint64_t k) {
    if (k < 2ull * G) return NodeHome{-1, k};
    const unsigned d = floor_log2_u64(k) - glog;  // depth below the subtree roots
    const uint64_t owner = (k >> d) - G;
    return NodeHome{(int)owner, (1ull << d) | (k & ((1ull << d) - 1))};
}

// all-to-all of packed per-row items: where rank `src`'s packed item t travels
inline uint32_t shard_leaf_destination(uint64_t leaves, uint32_t glog, uint64_t t) { return (uint32_t)(t / ((leaves >> glog) >> glog)); }
// ... and which local leaf of the destination it becomes (src = sending rank, t_in_chunk = t mod chunk)
inline uint64_t shard_local_leaf(uint32_t G, uint32_t src, uint64_t t_in_chunk) { return (uint64_t)G * t_in_chunk + src; }

}  // namespace ezk
