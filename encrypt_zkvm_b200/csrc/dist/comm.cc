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

void LocalGroup::send_recv(int rank, const void* send, void* recv, int src, size_t bytes, cudaStream_t s) {
    EZK_CUDA(cudaStreamSynchronize(s));
    {
        std::lock_guard<std::mutex> lock(mu_);
        slots_[rank] = send;
    }
    barrier();
    EZK_CUDA(cudaMemcpyAsync(recv, slots_[src], bytes, cudaMemcpyDefault, s));
    EZK_CUDA(cudaStreamSynchronize(s));
    barrier();
}

// ---------------------------------------------------------------------------------------------------------

void Comm::reset() {
    if (comm_) {
        api().CommDestroy(static_cast<ncclComm_t>(comm_));
        comm_ = nullptr;
    }
    local_ = nullptr;
    rank_ = 0, world_ = 1;
}

Comm::~Comm() {
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

void Comm::send_recv(const void* send, int dst, void* recv, int src, size_t bytes, cudaStream_t s) const {
    if (world_ == 1) {
        if (send != recv) EZK_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return;
    }
    if (local_) {
        local_->send_recv(rank_, send, recv, src, bytes, s);
        return;
    }
    ncclComm_t c = static_cast<ncclComm_t>(comm_);
    check(api().GroupStart(), "ncclGroupStart");
    check(api().Send(send, bytes, ncclUint8, dst, c, s), "ncclSend");
    check(api().Recv(recv, bytes, ncclUint8, src, c, s), "ncclRecv");
    check(api().GroupEnd(), "ncclGroupEnd");
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
