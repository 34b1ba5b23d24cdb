// Communicator of the coset-sharded single proof: one prover per GPU, exchanges over NVLink/NVSwitch.
//
// Two backends behind one interface:
//   * NCCL (one process per GPU, the production path): ncclAllGather, and grouped ncclSend/ncclRecv for the
//     all-to-all of leaf digests.  libnccl.so.2 is resolved with dlopen when a prover joins a group, so single-GPU
//     use carries no NCCL dependency (and a process that already imported torch shares torch's copy of the library).
//   * LocalGroup (several provers inside ONE process, each driven by its own host thread, on one or several GPUs):
//     host-synchronous device-to-device copies between the provers' buffers.  It exists so that the whole sharded
//     pipeline - ownership rules, packing orders, subtree Merkle commitments, opening routing - runs and is compared
//     byte for byte with the single-GPU proof on a ONE-GPU box (tests); it is not a performance path.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <cuda_runtime.h>

namespace ezk {

// Shared state of an in-process group.  Every member of the group calls the collectives in the same order.
class LocalGroup {
public:
    explicit LocalGroup(int world);
    int world() const { return world_; }
    // recv[q * bytes ..) = (all-gather) peer q's send[0 .. bytes) / (all-to-all) peer q's send[rank * bytes ..)
    void exchange(int rank, const void* send, void* recv, size_t bytes, bool all_to_all, cudaStream_t s);
    // a member that fails wakes the others instead of leaving them in the barrier
    void abort();
    // every member publishes a pointer; returns after all have, with all[q] = member q's pointer
    void publish(int rank, void* ptr, void* all[8]);
    void host_barrier() { barrier(); }

private:
    void barrier();
    const int world_;
    std::mutex mu_;
    std::condition_variable cv_;
    int waiting_ = 0;
    uint64_t generation_ = 0;
    bool aborted_ = false;
    const void* slots_[8] = {nullptr};
};

class Comm {
public:
    Comm() = default;
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;

    static void unique_id(uint8_t out[128]);                        // rank 0 creates it, every rank receives a copy
    void init(int rank, int world, const uint8_t id[128]);          // NCCL; collective over the `world` ranks
    void init_local(int rank, LocalGroup* group);                   // in-process group (group == nullptr: leave)
    bool active() const { return world_ > 1; }
    int rank() const { return rank_; }
    int world() const { return world_; }
    uint32_t world_log() const;
    // recv[q * bytes .. (q+1) * bytes) = rank q's send[0 .. bytes); send may be recv + rank * bytes (in place)
    void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t s) const;
    // recv[q * bytes .. (q+1) * bytes) = rank q's send[rank * bytes .. (rank+1) * bytes)
    void all_to_all(const void* send, void* recv, size_t bytes, cudaStream_t s) const;
    // called when a proof fails on this rank: in-process peers are released from their barriers
    void abort() const;

    // ---- peer memory (NVLink P2P stores fused into kernels) ----
    // Collective, once per proof: makes the workspaces of all ranks addressable from this one.  `base`/`bytes` = this
    // rank's workspace, `generation` changes whenever it is reallocated; `pinned`/`scratch` = small host / device staging
    // areas (>= 1 KiB).  NCCL ranks exchange cudaIpc handles (re-opened only when a generation changed); in-process
    // members exchange plain pointers.  Returns false - on every rank alike - when some mapping is not possible, in
    // which case the callers keep to the collectives above.
    bool map_peers(void* base, size_t bytes, uint64_t generation, void* pinned, void* scratch, cudaStream_t s);
    // rank q's copy of the address `local` of this rank's workspace (workspaces are laid out identically)
    template <class T>
    T* peer(int q, T* local) const {
        return reinterpret_cast<T*>(static_cast<char*>(peer_base_[q]) + (reinterpret_cast<char*>(local) - static_cast<char*>(own_base_)));
    }
    // all peer stores issued on the ranks' streams before this point are complete on every rank after it
    void barrier(void* scratch, cudaStream_t s) const;

private:
    void reset();
    void close_peers();
    void* own_base_ = nullptr;
    void* peer_base_[8] = {nullptr};
    uint64_t peer_gen_[8] = {0};
    bool peer_opened_[8] = {false};  // peer_base_[q] came from cudaIpcOpenMemHandle
    bool peers_ok_ = false;
    int rank_ = 0, world_ = 1;
    void* comm_ = nullptr;        // ncclComm_t
    LocalGroup* local_ = nullptr;  // not owned
};

}  // namespace ezk
