// mmrs_comm.hpp — NCCL binding of the library (one process per GPU, one communicator per context).
//
// The sweep path shards on independent work: whole units, or contiguous candidate sub-ranges of
// every unit, are dealt to ranks and only the tiny per-unit results cross NVLink — an all-reduce(MIN, uint64) of the
// packed (distance, index) keys when the candidate axis is split (the reduction of process_utils.rs:69-74: lowest
// distance, ties -> lowest index), an all-gather / all-reduce(SUM) of the 32-byte per-unit results otherwise. The
// collectives run on the context's stream on DEVICE buffers, between the kernels of a run: no host staging, no extra
// synchronisation.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded,
// elsewhere the system library), so libmmrs_b200.so has no link-time dependency on it and still loads where NCCL is
// absent; only mmrs_comm_unique_id / mmrs_ctx_comm_init fail there.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <mutex>
#include <string>

namespace mmrs {

// The subset of nccl.h this library uses (NCCL 2.x ABI; the values are part of NCCL's public, stable interface).
using nccl_comm_t = void*;
struct nccl_unique_id {
    char internal[128];
};
enum : int { kNcclSuccess = 0 };
enum : int { kNcclInt32 = 2, kNcclInt64 = 4, kNcclUint64 = 5 };   // ncclDataType_t
enum : int { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };           // ncclRedOp_t

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool ok() const { return handle && GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && Broadcast; }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed");
            return;
        }
        auto sym = [&](const char* s) { return dlsym(api.handle, s); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!api.ok()) api.error = "NCCL is not available: libnccl.so.2 lacks an expected symbol";
    });
    return api;
}

inline std::string nccl_err(int rc) {
    NcclApi& a = nccl_api();
    return std::string("NCCL: ") + (a.GetErrorString ? a.GetErrorString(rc) : "error ") + " (" + std::to_string(rc) + ")";
}

}  // namespace mmrs
