// See comm.h.
#include "comm.h"
#include "../common.h"
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
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return a;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw CudaError(std::string(what) + " failed: " + api().GetErrorString(r));
}

}  // namespace

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
    if (comm_) {
        api().CommDestroy(static_cast<ncclComm_t>(comm_));
        comm_ = nullptr;
    }
    rank_ = 0, world_ = 1;  // a failed initialisation below leaves a working single-GPU prover
    if (world == 1) return;
    ncclUniqueId id;
    memcpy(&id, idb, 128);
    ncclComm_t c = nullptr;
    check(api().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
    comm_ = c;
    rank_ = rank, world_ = world;
}

uint32_t Comm::world_log() const { return world_ == 8 ? 3 : world_ == 4 ? 2 : world_ == 2 ? 1 : 0; }

void Comm::all_gather(const void* send, void* recv, size_t bytes, cudaStream_t s) const {
    if (world_ == 1) {
        if (send != recv) EZK_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return;
    }
    check(api().AllGather(send, recv, bytes, ncclUint8, static_cast<ncclComm_t>(comm_), s), "ncclAllGather");
}

}  // namespace ezk
