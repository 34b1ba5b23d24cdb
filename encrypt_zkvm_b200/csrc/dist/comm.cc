// See comm.h.
#include "comm.h"
#include "../common.h"
#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <string>

namespace ezk {

namespace {

struct Api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // torch's copy when it is already in the process
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw CudaError(std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* name) {
            void* p = dlsym(h, name);
            if (!p) throw CudaError(std::string("libnccl is missing ") + name);
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return a;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw CudaError(std::string(what) + " failed: " + api().GetErrorString(r));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// in-process group

LocalGroup::LocalGroup(int world) : world_(world) {
    if (world != 1 && world != 2 && world != 4 && world != 8) throw CudaError("the prover shards over 1, 2, 4 or 8 members");
}

void LocalGroup::abort() {
    {
        std::lock_guard<std::mutex> lock(mu_);
        aborted_ = true;
    }
    cv_.notify_all();
}

void LocalGroup::barrier() {
    std::unique_lock<std::mutex> lock(mu_);
    if (aborted_) throw CudaError("another member of the in-process group failed");
    const uint64_t gen = generation_;
    if (++waiting_ == world_) {
        waiting_ = 0;
        generation_++;
        cv_.notify_all();
        return;
    }
    const bool ok = cv_.wait_for(lock, std::chrono::seconds(120), [&] { return generation_ != gen || aborted_; });
    if (aborted_) throw CudaError("another member of the in-process group failed");
    if (!ok) {
        aborted_ = true;
        cv_.notify_all();
        throw CudaError("in-process group: a member did not reach the collective within 120 s");
    }
}

void LocalGroup::publish(int rank, void* ptr, void* all[8]) {
    {
        std::lock_guard<std::mutex> lock(mu_);
        slots_[rank] = ptr;
    }
    barrier();
    for (int q = 0; q < world_; q++) all[q] = const_cast<void*>(slots_[q]);
    barrier();
}

void LocalGroup::exchange(int rank, const void* send, void* recv, size_t bytes, bool all_to_all, cudaStream_t s) {
    EZK_CUDA(cudaStreamSynchronize(s));  // this member's data is complete
    {
        std::lock_guard<std::mutex> lock(mu_);
        slots_[rank] = send;
    }
    barrier();
    for (int q = 0; q < world_; q++) {
        const uint8_t* src = static_cast<const uint8_t*>(slots_[q]) + (all_to_all ? (size_t)rank * bytes : 0);
        uint8_t* dst = static_cast<uint8_t*>(recv) + (size_t)q * bytes;
        if (src != dst) EZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
    }
    EZK_CUDA(cudaStreamSynchronize(s));
    barrier();  // nobody reuses its send buffer before every member has read it
}

// ---------------------------------------------------------------------------------------------------------

void Comm::close_peers() {
    for (int q = 0; q < 8; q++) {
        if (peer_opened_[q] && peer_base_[q]) cudaIpcCloseMemHandle(peer_base_[q]);
        peer_base_[q] = nullptr, peer_gen_[q] = 0, peer_opened_[q] = false;
    }
    peers_ok_ = false, own_base_ = nullptr;
}

namespace {
struct PeerRecord {  // what every rank publishes about its workspace
    cudaIpcMemHandle_t handle;
    uint64_t generation;
    uint64_t bytes;
    uint32_t ok;  // second round: all mappings of this rank are in place
    uint32_t pad;
};
}  // namespace

bool Comm::map_peers(void* base, size_t bytes, uint64_t generation, void* pinned, void* scratch, cudaStream_t s) {
    if (world_ == 1) return false;
    own_base_ = base;
    if (local_) {
        // members of one process: plain pointers (peer access between their devices is enabled on demand)
        void* all[8];
        local_->publish(rank_, base, all);
        int dev = 0, ok = 1;
        EZK_CUDA(cudaGetDevice(&dev));
        for (int q = 0; q < world_; q++) {
            peer_base_[q] = all[q], peer_opened_[q] = false;
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, all[q]) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else if (attr.device != dev) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                cudaGetLastError();
            }
        }
        void* oks[8];
        local_->publish(rank_, ok ? base : nullptr, oks);
        peers_ok_ = true;
        for (int q = 0; q < world_; q++) peers_ok_ = peers_ok_ && oks[q] != nullptr;
        return peers_ok_;
    }
    static_assert(sizeof(PeerRecord) == 88, "PeerRecord layout");
    PeerRecord* host = static_cast<PeerRecord*>(pinned);
    PeerRecord* dev = static_cast<PeerRecord*>(scratch);
    auto gather = [&](const PeerRecord& mine, PeerRecord out[8]) {
        host[0] = mine;
        EZK_CUDA(cudaMemcpyAsync(dev, host, sizeof(PeerRecord), cudaMemcpyHostToDevice, s));
        check(api().AllGather(dev, dev + 1, sizeof(PeerRecord), ncclUint8, static_cast<ncclComm_t>(comm_), s), "ncclAllGather");
        EZK_CUDA(cudaMemcpyAsync(host + 1, dev + 1, sizeof(PeerRecord) * world_, cudaMemcpyDeviceToHost, s));
        EZK_CUDA(cudaStreamSynchronize(s));
        memcpy(out, host + 1, sizeof(PeerRecord) * world_);
    };
    PeerRecord mine{};
    if (cudaIpcGetMemHandle(&mine.handle, base) != cudaSuccess) {
        cudaGetLastError();
        memset(&mine.handle, 0, sizeof(mine.handle));
        mine.generation = 0;  // 0 = no handle: every rank falls back
    } else {
        mine.generation = generation ? generation : 1;
    }
    mine.bytes = bytes;
    PeerRecord all[8];
    gather(mine, all);
    bool changed = false, possible = true;
    for (int q = 0; q < world_; q++) {
        possible = possible && all[q].generation != 0;
        changed = changed || all[q].generation != peer_gen_[q];
    }
    if (!possible) {
        close_peers();
        own_base_ = base;
        return false;
    }
    if (!changed) return peers_ok_;
    uint32_t ok = 1;
    for (int q = 0; q < world_; q++) {
        if (all[q].generation == peer_gen_[q]) continue;
        if (peer_opened_[q] && peer_base_[q]) cudaIpcCloseMemHandle(peer_base_[q]);
        peer_base_[q] = nullptr, peer_opened_[q] = false, peer_gen_[q] = 0;
        if (q == rank_) {
            peer_base_[q] = base, peer_gen_[q] = all[q].generation;
            continue;
        }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            continue;
        }
        peer_base_[q] = p, peer_opened_[q] = true, peer_gen_[q] = all[q].generation;
    }
    mine.ok = ok;
    gather(mine, all);  // the same verdict on every rank
    peers_ok_ = true;
    for (int q = 0; q < world_; q++) peers_ok_ = peers_ok_ && all[q].ok != 0;
    return peers_ok_;
}

void Comm::barrier(void* scratch, cudaStream_t s) const {
    if (world_ == 1) return;
    if (local_) {
        EZK_CUDA(cudaStreamSynchronize(s));
        local_->host_barrier();
        return;
    }
    // a 4-byte all-gather on the stream: it completes on a rank only after every rank has reached it, i.e. after the
    // kernels those ranks queued before it - and their peer stores - are done
    uint32_t* d = static_cast<uint32_t*>(scratch);
    check(api().AllGather(d + rank_, d, sizeof(uint32_t), ncclUint8, static_cast<ncclComm_t>(comm_), s), "ncclAllGather");
}

void Comm::reset() {
    close_peers();
    if (comm_) {
        api().CommDestroy(static_cast<ncclComm_t>(comm_));
        comm_ = nullptr;
    }
    local_ = nullptr;
    rank_ = 0, world_ = 1;
}

Comm::~Comm() {
    close_peers();
    if (comm_) api().CommDestroy(static_cast<ncclComm_t>(comm_));
}

void Comm::unique_id(uint8_t out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, &id, 128);
}

void Comm::init(int rank, int world, const uint8_t idb[128]) {
    if (world != 1 && world != 2 && world != 4 && world != 8) throw CudaError("the prover shards over 1, 2, 4 or 8 GPUs");
    if (rank < 0 || rank >= world) throw CudaError("bad rank");
    reset();  // a failed initialisation below leaves a working single-GPU prover
    if (world == 1) return;
    ncclUniqueId id;
    memcpy(&id, idb, 128);
    ncclComm_t c = nullptr;
    check(api().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
    comm_ = c;
    rank_ = rank, world_ = world;
}

void Comm::init_local(int rank, LocalGroup* group) {
    reset();
    if (!group || group->world() == 1) return;
    if (rank < 0 || rank >= group->world()) throw CudaError("bad rank");
    local_ = group;
    rank_ = rank, world_ = group->world();
}

uint32_t Comm::world_log() const { return world_ == 8 ? 3 : world_ == 4 ? 2 : world_ == 2 ? 1 : 0; }

void Comm::abort() const {
    if (local_) local_->abort();
}

void Comm::all_gather(const void* send, void* recv, size_t bytes, cudaStream_t s) const {
    if (world_ == 1) {
        if (send != recv) EZK_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return;
    }
    if (local_) {
        local_->exchange(rank_, send, recv, bytes, false, s);
        return;
    }
    check(api().AllGather(send, recv, bytes, ncclUint8, static_cast<ncclComm_t>(comm_), s), "ncclAllGather");
}

void Comm::all_to_all(const void* send, void* recv, size_t bytes, cudaStream_t s) const {
    if (world_ == 1) {
        if (send != recv) EZK_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return;
    }
    if (local_) {
        local_->exchange(rank_, send, recv, bytes, true, s);
        return;
    }
    ncclComm_t c = static_cast<ncclComm_t>(comm_);
    check(api().GroupStart(), "ncclGroupStart");
    for (int q = 0; q < world_; q++) {
        check(api().Send(static_cast<const uint8_t*>(send) + (size_t)q * bytes, bytes, ncclUint8, q, c, s), "ncclSend");
        check(api().Recv(static_cast<uint8_t*>(recv) + (size_t)q * bytes, bytes, ncclUint8, q, c, s), "ncclRecv");
    }
    check(api().GroupEnd(), "ncclGroupEnd");
}

}  // namespace ezk
