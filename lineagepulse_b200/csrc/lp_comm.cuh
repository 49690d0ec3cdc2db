// lp_comm.cuh -- NCCL inside the library.  The cross-shard sums of the gene-sharded path (per-cell drop-out
// partial sums of every lock-step fitPi round, the EM log-likelihood, convergence / error flags) are all-reduced
// by NCCL over NVLink on the context's own stream: no host callback in the hot loop.
//
// NCCL is bound at run time (dlopen), not at link time: liblpb200.so must load on a box without NCCL and, inside
// a Python process that also imports torch, must share torch's copy (two NCCL copies with the same SONAME in one
// process resolve to whichever was loaded first).  Search order: $LPB200_NCCL_LIBRARY, libnccl.so.2, libnccl.so.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[3] = {getenv("LPB200_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
            api.error = dlerror();
        }
        if (!api.handle) return;
#define LP_NCCL_SYM(field, name)                                            \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return; }
        LP_NCCL_SYM(GetVersion, "ncclGetVersion")
        LP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        LP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        LP_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        LP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        LP_NCCL_SYM(AllReduce, "ncclAllReduce")
        LP_NCCL_SYM(AllGather, "ncclAllGather")
        LP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LP_NCCL_SYM
    });
    return api.handle ? &api : nullptr;
}
static const char *nccl_load_error() {
    static NcclApi *dummy = nccl_api();
    (void)dummy;
    static std::string msg;
    NcclApi *a = nccl_api();
    if (a) return "";
    // the static inside nccl_api() keeps the dlerror text
    msg = "NCCL library not found (set LPB200_NCCL_LIBRARY)";
    return msg.c_str();
}
