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
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
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
        LP_NCCL_SYM(GroupStart, "ncclGroupStart")
        LP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        LP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LP_NCCL_SYM
    });
    return api.handle ? &api : nullptr;
}

// ncclGroupStart / ncclGroupEnd around collectives that one thread issues for several devices
struct NcclGroup {
    NcclApi *api;
    explicit NcclGroup(NcclApi *a) : api(a) { api->GroupStart(); }
    ~NcclGroup() { api->GroupEnd(); }
};

// Loopback communicator: several shard contexts on ONE device (lpb200_create_multi with a repeated device id).
// NCCL cannot put two ranks on one GPU; here the shards' host threads meet at a barrier and every shard adds up
// all the shards' buffers (same device, same address space) in rank order.  Exists so that the multi-device
// machinery -- slicing, gathering, the collective sequence of the EM loop -- runs and is tested on a single-GPU
// box exactly as it does over NCCL; there is no reason to use it in production.
#include <condition_variable>
struct Loopback {
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long generation = 0;
    std::vector<const void *> bufs;
    std::vector<void *> tmp;
    std::vector<size_t> tmp_words;
    bool broken = false;
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long gen = generation;
        if (++waiting == n) {
            waiting = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen || broken; });
        }
    }
};

__global__ void loopback_sum_kernel(const void *b0, const void *b1, const void *b2, const void *b3, const void *b4, const void *b5,
                                    const void *b6, const void *b7, int n, void *out, size_t count, int is_f64) {
    const void *b[8] = {b0, b1, b2, b3, b4, b5, b6, b7};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        if (is_f64) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s += reinterpret_cast<const double *>(b[k])[i];
            reinterpret_cast<double *>(out)[i] = s;
        } else {
            long long s = 0;
            for (int k = 0; k < n; k++) s += reinterpret_cast<const long long *>(b[k])[i];
            reinterpret_cast<long long *>(out)[i] = s;
        }
    }
}

static int loopback_allreduce(Loopback *L, int rank, void *dev, size_t count, int is_f64, cudaStream_t stream) {
    if (L->n > 8) return 1;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 2;
    L->bufs[rank] = dev;
    if (L->tmp_words[rank] < count) {
        if (L->tmp[rank]) cudaFree(L->tmp[rank]);
        if (cudaMalloc(&L->tmp[rank], 8 * count) != cudaSuccess) return 3;
        L->tmp_words[rank] = count;
    }
    L->barrier();   // every shard's buffer is complete and registered
    const void *b[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < L->n; k++) b[k] = L->bufs[k];
    loopback_sum_kernel<<<64, 256, 0, stream>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], L->n, L->tmp[rank], count, is_f64);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 4;
    L->barrier();   // every shard has read every buffer
    if (cudaMemcpyAsync(dev, L->tmp[rank], 8 * count, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) return 5;
    return 0;
}
