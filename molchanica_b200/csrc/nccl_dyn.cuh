// nccl_dyn.cuh -- NCCL resolved at run time (dlopen), the way the reference resolves the CUDA driver
// (cudarc "dynamic-loading", reference Cargo.toml:56-65): a single-GPU host never needs libnccl, and
// a process that already carries an NCCL (PyTorch bundles its own) keeps using exactly that copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <string>

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
    std::string err;
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    if (api.ok || !api.err.empty()) return api;
    // MOLCHANICA_NCCL_LIB names the library explicitly (the host build of tests/cpp/host_lib/ points it at a
    // shared-memory stand-in with the same ten entry points)
    const char *forced = getenv("MOLCHANICA_NCCL_LIB");
    void *h = forced ? dlopen(forced, RTLD_NOW | RTLD_GLOBAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // else: an NCCL this process already loaded
    if (!h && !forced) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h && !forced) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
#define MC_SYM(field, name)                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                   \
    if (!api.field) { api.err = std::string("libnccl lacks ") + name; return api; }
    MC_SYM(GetUniqueId, "ncclGetUniqueId")
    MC_SYM(CommInitRank, "ncclCommInitRank")
    MC_SYM(CommDestroy, "ncclCommDestroy")
    MC_SYM(GetErrorString, "ncclGetErrorString")
    MC_SYM(GroupStart, "ncclGroupStart")
    MC_SYM(GroupEnd, "ncclGroupEnd")
    MC_SYM(Send, "ncclSend")
    MC_SYM(Recv, "ncclRecv")
    MC_SYM(AllGather, "ncclAllGather")
    MC_SYM(AllReduce, "ncclAllReduce")
#undef MC_SYM
    api.ok = true;
    return api;
}
