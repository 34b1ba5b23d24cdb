// NCCL communicator for the coset-sharded single proof: one process per GPU, NVLink/NVSwitch
// all-gathers of the small per-row products (leaf digests, constraint evaluations, DEEP evaluations, opened rows).
// libnccl.so.2 is resolved with dlopen when a prover joins a group, so single-GPU use carries no NCCL dependency
// (and a process that already imported torch shares torch's copy of the library).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace ezk {

class Comm {
public:
    Comm() = default;
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;

    static void unique_id(uint8_t out[128]);                        // rank 0 creates it, every rank receives a copy
    void init(int rank, int world, const uint8_t id[128]);          // collective over the `world` ranks
    bool active() const { return world_ > 1; }
    int rank() const { return rank_; }
    int world() const { return world_; }
    uint32_t world_log() const;
    // recv[q * bytes .. (q+1) * bytes) = rank q's send[0 .. bytes)
    void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t s) const;

private:
    int rank_ = 0, world_ = 1;
    void* comm_ = nullptr;  // ncclComm_t
};

}  // namespace ezk
