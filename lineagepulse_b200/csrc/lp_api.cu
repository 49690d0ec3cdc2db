// lp_api.cu -- the C ABI of include/lpb200.h: context, data loading, host drivers of the kernels
// (fitMuDisp, fitPi, evalLogLikMatrix, fitModel's EM loop) and the LRT.  No CPU fallback: every
// compute entry point needs a CUDA device and fails with LPB200_E_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "lp_kernels.cuh"
#include "lp_comm.cuh"

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct lpb200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    // counts
    int G = 0, N = 0;
    bool is_u16 = false;
    void *d_counts = nullptr;  // uint16_t or int32_t, gene-major
    size_t pitch = 0;
    int32_t *d_xmax = nullptr;
    uint32_t *d_perm = nullptr;   // [G][pitch] non-zero cells first, then zeros
    int32_t *d_nnz = nullptr, *d_nobs = nullptr;
    std::vector<double> row_means;
    std::vector<int32_t> row_max, col_max;   // per gene / per cell largest count (all-zero filters of processSCData)
    unsigned long long n_negative = 0;       // entries < 0 (NA markers included)
    int global_max = 0;
    // design
    bool have_design = false;
    DevDesign D{};
    std::vector<void *> design_allocs;
    std::vector<double> h_t;
    double *d_Q = nullptr;        // orthonormal basis of impulse_basis, N x 4
    bool have_starts = false;
    std::vector<double> h_peak, h_valley;  // G x 7 column-major
    std::vector<double> h_bs_mu;           // host copy of the mu spline basis (for Pinv)
    // collective: the gene shards of one matrix.  Either an NCCL communicator owned by the library
    // (lpb200_comm_init: one process per GPU; lpb200_create_multi: one process, one shard context per device)
    // or the host's all-reduce callback (lpb200_set_collective).
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    bool comm_borrowed = false;     // the communicator belongs to the process-wide cache (lp_multi.cuh)
    struct CommSet *comm_set = nullptr;
    Loopback *loopback = nullptr;   // shards of a multi-device context that share one device (testing)
    lpb200_allreduce_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    double *d_xchg = nullptr;  // small exchange buffer
    uint64_t allreduce_calls = 0, allreduce_bytes = 0;
    // multi-device context (lpb200_create_multi): this object only dispatches to its shard contexts
    std::vector<lpb200_ctx *> shards;
    std::vector<int> bounds;       // gene range of shard k: [bounds[k], bounds[k + 1])
    bool owns_stream = false;
    // stats
    lpb200_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaStream_t> aux_streams;   // for concurrent independent fits (lpb200_fit_model_multi)
    int use_table = 1;
    int no_fastpath = 0;   // LPB200_NO_FASTPATH=1: always run the generic evaluator (testing)
    int ctas_per_sm = LP_FAST_MINBLOCKS;  // persistent fit CTAs per SM (LPB200_CTAS_PER_SM overrides)
    int cluster = 0;       // CTAs per work item of the fit kernel; 0: chosen per launch (LPB200_CLUSTER = 1 | 2 overrides)
};

// The fit-kernel variants: count type (uint16_t / int32_t) x evaluator (generic; constant or impulse mean x NB or ZINB;
// groups or splines mean x NB or ZINB) x launched as thread-block clusters or not.  They are instantiated in two
// translation units of their own (lp_fit_u16.cu, lp_fit_i32.cu: see lp_fit_table.cuh) so that the library builds in
// parallel.
#include "lp_fit_table.cuh"
static FitKernelFn fit_kernel(bool u16, int variant, bool cl) {
    return u16 ? lp_fit_kernel_u16(variant, cl) : lp_fit_kernel_i32(variant, cl);
}
static size_t fit_smem_bytes(bool cl) { return cl ? sizeof(FitSharedT<true>) : sizeof(FitSharedT<false>); }

// multi-device contexts (lp_multi.cuh, included at the end of this file)
static bool is_multi(const lpb200_ctx *c);
static int multi_load_counts_dense(lpb200_ctx *ctx, int G, int N, const double *x);
static int multi_load_counts_csc(lpb200_ctx *ctx, int G, int N, const int32_t *p, const int32_t *i, const double *x);
static int multi_load_counts_i32(lpb200_ctx *ctx, int G, int N, const int32_t *x, int is_device);
static int multi_set_design(lpb200_ctx *ctx, const lpb200_design *d);
static int multi_row_means(lpb200_ctx *ctx, double *means);
static int multi_init_params(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *out, double min_rowmean_global);
static int multi_eval_loglik(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, double *ll);
static int multi_impulse_starts(lpb200_ctx *ctx, double *peak, double *valley);
static int multi_fit_mudisp_items(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll,
                                  int32_t *conv, int32_t *n_starts, int32_t *n_theta, double *theta0, double *theta,
                                  double *ll_items, int32_t *conv_items, uint32_t *evals_items);
static int multi_fit_pi(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll, int32_t *conv);
static int multi_fit_model(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, int fit_drop, const lpb200_control *c,
                           double min_rowmean_global, int keep_mask, double *em_ll, int32_t *n_iter, int32_t *converged,
                           double *ll_init, double *ll, int32_t *conv);
static int multi_fit_model_multi(lpb200_ctx *ctx, int n_jobs, const lpb200_model *models, lpb200_params *params,
                                 const lpb200_control *c, double min_rowmean_global, double *ll_init, double *em_ll,
                                 int32_t *converged, double *ll, int32_t *conv, uint64_t *ole);
static int multi_expand_fits(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, const int32_t *gene_ids, int n_ids,
                             int what, double *z);
static int multi_get_stats(lpb200_ctx *ctx, lpb200_stats *s);
static void release_comms(struct CommSet *cs);

// LPB200_DEBUG: host wall-clock of the phases of a call
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *scope;
    explicit PhaseTimer(const char *scope_) : on(getenv("LPB200_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()), scope(scope_) {}
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[lpb200] %s: %s %.2f ms\n", scope, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static int set_err(lpb200_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return set_err(ctx, LPB200_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Scratch-buffer pool: cudaMalloc / cudaFree synchronise the device and showed up as 0.3-1 s of
// host-side jitter per fit; freed scratch blocks are kept per device and reused by later calls.
struct PoolBlock {
    void *p;
    size_t bytes;
    int device;
};
static std::vector<PoolBlock> g_pool;
static size_t g_pool_bytes = 0;
static std::mutex g_pool_mutex;   // contexts of several host threads share the pool
static const size_t POOL_CAP_BYTES = (size_t)32 << 30;   // over all devices of the process

static cudaError_t pool_alloc(void **out, size_t bytes, size_t *got) {
    int dev = 0;
    cudaGetDevice(&dev);
    bytes = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    std::unique_lock<std::mutex> lock(g_pool_mutex);
    size_t best = (size_t)-1;
    for (size_t k = 0; k < g_pool.size(); k++)
        if (g_pool[k].device == dev && g_pool[k].bytes >= bytes && g_pool[k].bytes <= 2 * bytes + ((size_t)1 << 20) &&
            (best == (size_t)-1 || g_pool[k].bytes < g_pool[best].bytes))
            best = k;
    if (best != (size_t)-1) {
        *out = g_pool[best].p;
        *got = g_pool[best].bytes;
        g_pool_bytes -= g_pool[best].bytes;
        g_pool.erase(g_pool.begin() + best);
        return cudaSuccess;
    }
    lock.unlock();
    *got = bytes;
    return cudaMalloc(out, bytes);
}
static void pool_free(void *p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (g_pool_bytes + bytes <= POOL_CAP_BYTES) {
            g_pool.push_back(PoolBlock{p, bytes, dev});
            g_pool_bytes += bytes;
            return;
        }
    }
    cudaFree(p);
}
static void pool_release(int device) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (size_t k = 0; k < g_pool.size();) {
        if (g_pool[k].device == device) {
            cudaFree(g_pool[k].p);
            g_pool_bytes -= g_pool[k].bytes;
            g_pool.erase(g_pool.begin() + k);
        } else {
            k++;
        }
    }
}

template <typename T>
struct DevBuf {  // RAII device scratch buffer (pooled)
    T *p = nullptr;
    size_t n = 0, bytes = 0;
    DevBuf() {}
    ~DevBuf() { if (p) pool_free(p, bytes); }
    cudaError_t alloc(size_t count) {
        n = count;
        return pool_alloc((void **)&p, count * sizeof(T), &bytes);
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

extern "C" int lpb200_abi_version(void) { return LPB200_ABI_VERSION; }

extern "C" void lpb200_default_control(lpb200_control *c) {
    c->maxit_mudisp = 1000;
    c->reltol_mudisp = 1e-8;
    c->maxit_pi = 10000;
    c->reltol_pi = 1e-8;
    c->ndeps = 1e-3;
    c->max_cycles = 20;
    c->prec_em = 1.0 - 1e-4;
}

extern "C" int lpb200_n_mu_params(const lpb200_model *m, const lpb200_design *d) {
    return lp_n_mu_params(m->mu_model, d->k_spline_mu, d->n_groups);
}
extern "C" int lpb200_n_disp_params(const lpb200_model *m, const lpb200_design *d) {
    return lp_n_disp_params(m->disp_model, d->k_spline_disp, d->n_groups);
}
extern "C" int lpb200_n_drop_params(const lpb200_model *m, const lpb200_design *d) {
    if (m->drop_model == LPB200_DROP_NONE) return 0;
    return 1 + (m->drop_model == LPB200_DROP_LOGISTIC_OFMU ? 1 : 0) + d->n_pi_pred;
}
extern "C" int lpb200_n_batch_cols(const int32_t *nb, int n_conf) {
    int s = 0;
    for (int c = 0; c < n_conf; c++) s += nb[c];
    return s;
}
extern "C" int lpb200_deg_freedom(const lpb200_model *m, const lpb200_design *d) {
    return lpb200_n_mu_params(m, d) + lpb200_n_batch_cols(d->n_batches_mu, d->n_conf_mu) - d->n_conf_mu +
           lpb200_n_disp_params(m, d) + lpb200_n_batch_cols(d->n_batches_disp, d->n_conf_disp) - d->n_conf_disp;
}

extern "C" lpb200_ctx *lpb200_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    lpb200_ctx *c = new lpb200_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaMalloc((void **)&c->d_xchg, 64 * sizeof(double)) != cudaSuccess) {
        delete c;
        return nullptr;
    }
    // load every fit-kernel variant now: with CUDA's lazy module loading the first launch of a kernel can
    // wait for running kernels, which would serialise the concurrent fits of lpb200_fit_model_multi.  Once per
    // device and process: the primary context keeps the modules, later contexts skip the ~20 ms.
    {
        static std::mutex loaded_mutex;
        static std::vector<int> loaded_devices;
        std::lock_guard<std::mutex> lock(loaded_mutex);
        if (std::find(loaded_devices.begin(), loaded_devices.end(), device) == loaded_devices.end()) {
            for (int ct = 0; ct < 2; ct++)
                for (int cl = 0; cl < 2; cl++)
                    for (int v = 0; v < LP_N_FIT_VARIANTS; v++) {
                        cudaFuncSetAttribute((const void *)fit_kernel(ct != 0, v, cl != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)fit_smem_bytes(cl != 0));
                        cudaFuncAttributes fa;
                        cudaFuncGetAttributes(&fa, (const void *)fit_kernel(ct != 0, v, cl != 0));   // forces the module load
                    }
            cudaGetLastError();
            loaded_devices.push_back(device);
        }
    }
    const char *e = getenv("LPB200_NO_TABLE");
    if (e && atoi(e)) c->use_table = 0;
    e = getenv("LPB200_CTAS_PER_SM");
    if (e && atoi(e) > 0) c->ctas_per_sm = atoi(e);
    e = getenv("LPB200_NO_FASTPATH");
    if (e && atoi(e)) c->no_fastpath = 1;
    e = getenv("LPB200_CLUSTER");
    if (e && (atoi(e) == 1 || atoi(e) == 2)) c->cluster = atoi(e);
    return c;
}

static void free_design(lpb200_ctx *c) {
    for (void *p : c->design_allocs) cudaFree(p);
    c->design_allocs.clear();
    c->d_Q = nullptr;
    c->have_design = false;
    c->have_starts = false;
}

extern "C" void lpb200_destroy(lpb200_ctx *c) {
    if (!c) return;
    if (is_multi(c)) {
        Loopback *L = c->shards[0]->loopback;
        for (lpb200_ctx *s : c->shards) lpb200_destroy(s);
        release_comms(c->comm_set);
        if (L) {
            for (void *t : L->tmp) if (t) cudaFree(t);
            delete L;
        }
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    // the scratch pool outlives the context (a caching allocator: the public API makes one context per count matrix,
    // and a fresh pool would cost every new context some dozens of cudaMalloc); lpb200_release_cached_memory frees it
    free_design(c);
    if (c->d_counts) cudaFree(c->d_counts);
    if (c->d_xmax) cudaFree(c->d_xmax);
    if (c->d_perm) cudaFree(c->d_perm);
    if (c->d_nnz) cudaFree(c->d_nnz);
    if (c->d_nobs) cudaFree(c->d_nobs);
    if (c->d_xchg) cudaFree(c->d_xchg);
    for (cudaStream_t st : c->aux_streams) cudaStreamDestroy(st);
    if (c->comm && !c->comm_borrowed && nccl_api()) nccl_api()->CommDestroy(c->comm);
    if (c->owns_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
}

extern "C" int lpb200_release_cached_memory(int device) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return LPB200_E_CUDA;
    pool_release(device);
    cudaSetDevice(prev);
    return 0;
}

extern "C" const char *lpb200_last_error(lpb200_ctx *c) { return c ? c->err.c_str() : "no context (CUDA device unavailable)"; }

extern "C" int lpb200_set_stream(lpb200_ctx *ctx, void *s) {
    if (!ctx) return LPB200_E_ARG;
    if (is_multi(ctx)) return set_err(ctx, LPB200_E_ARG, "a multi-device context owns one stream per device");
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->owns_stream = false;
    ctx->stream = (cudaStream_t)s;
    return 0;
}

extern "C" int lpb200_set_collective(lpb200_ctx *ctx, int rank, int world, lpb200_allreduce_fn fn, void *user) {
    if (is_multi(ctx)) return set_err(ctx, LPB200_E_ARG, "a multi-device context has its own communicator");
    if (!ctx || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return set_err(ctx, LPB200_E_ARG, "bad collective");
    if (ctx->comm && nccl_api()) {
        nccl_api()->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->rank = rank;
    ctx->world = world;
    ctx->allreduce = fn;
    ctx->allreduce_user = user;
    return 0;
}

extern "C" int lpb200_get_stats(lpb200_ctx *ctx, lpb200_stats *s) {
    if (!ctx || !s) return LPB200_E_ARG;
    if (is_multi(ctx)) return multi_get_stats(ctx, s);
    *s = ctx->stats;
    return 0;
}
extern "C" int lpb200_reset_stats(lpb200_ctx *ctx) {
    if (!ctx) return LPB200_E_ARG;
    for (lpb200_ctx *c : ctx->shards) lpb200_reset_stats(c);
    ctx->stats = lpb200_stats{};
    ctx->allreduce_calls = ctx->allreduce_bytes = 0;
    return 0;
}

// ---- NCCL communicator of a one-process-per-GPU run -----------------------------------------------------------
extern "C" int lpb200_comm_unique_id(char *id) {
    NcclApi *nccl = nccl_api();
    if (!nccl || !id) return LPB200_E_COMM;
    ncclUniqueId u;
    if (nccl->GetUniqueId(&u) != ncclSuccess) return LPB200_E_COMM;
    memcpy(id, u.internal, LPB200_UNIQUE_ID_BYTES);
    return 0;
}

extern "C" int lpb200_comm_init(lpb200_ctx *ctx, int rank, int world, const char *id) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) return set_err(ctx, LPB200_E_ARG, "a multi-device context has its own communicator");
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) return set_err(ctx, LPB200_E_ARG, "comm_init: bad arguments");
    NcclApi *nccl = nccl_api();
    if (ctx->comm && nccl) {
        nccl->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->rank = rank;
    ctx->world = world;
    ctx->allreduce = nullptr;
    if (world == 1) return 0;
    if (!nccl) return set_err(ctx, LPB200_E_COMM, "NCCL library not found (set LPB200_NCCL_LIBRARY)");
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(u.internal, id, LPB200_UNIQUE_ID_BYTES);
    ncclResult_t r = nccl->CommInitRank(&ctx->comm, world, u, rank);
    if (r != ncclSuccess) {
        ctx->comm = nullptr;
        ctx->world = 1;
        ctx->rank = 0;
        return set_err(ctx, LPB200_E_COMM, "ncclCommInitRank failed: %s", nccl->GetErrorString(r));
    }
    // NCCL sets a communicator's connections up at its first collective (0.3-1.5 s for 2-8 ranks): do that here, so
    // that the first all-reduce of a fit is an all-reduce
    CK(cudaMemsetAsync(ctx->d_xchg + 16, 0, sizeof(double), ctx->stream));
    r = nccl->AllReduce(ctx->d_xchg + 16, ctx->d_xchg + 16, 1, ncclFloat64, ncclSum, ctx->comm, ctx->stream);
    if (r != ncclSuccess) return set_err(ctx, LPB200_E_COMM, "ncclAllReduce (warm-up) failed: %s", nccl->GetErrorString(r));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lpb200_comm_stats(lpb200_ctx *ctx, uint64_t *calls, uint64_t *bytes) {
    if (!ctx || !calls || !bytes) return LPB200_E_ARG;
    const lpb200_ctx *c = is_multi(ctx) ? ctx->shards[0] : ctx;
    *calls = c->allreduce_calls;
    *bytes = c->allreduce_bytes;
    return 0;
}

static int do_allreduce(lpb200_ctx *ctx, void *dev, size_t count, int dtype) {
    if (ctx->world <= 1) return 0;
    ctx->allreduce_calls++;
    ctx->allreduce_bytes += 8 * (uint64_t)count;
    if (ctx->comm) {   // stream-ordered, in place
        ncclResult_t r = nccl_api()->AllReduce(dev, dev, count, dtype == LPB200_I64 ? ncclInt64 : ncclFloat64, ncclSum, ctx->comm, ctx->stream);
        if (r != ncclSuccess) return set_err(ctx, LPB200_E_COMM, "ncclAllReduce failed: %s", nccl_api()->GetErrorString(r));
        return 0;
    }
    if (ctx->loopback) {
        int rc = loopback_allreduce(ctx->loopback, ctx->rank, dev, count, dtype == LPB200_F64, ctx->stream);
        if (rc != 0) return set_err(ctx, LPB200_E_COMM, "loopback all-reduce failed (%d)", rc);
        return 0;
    }
    int rc = ctx->allreduce(ctx->allreduce_user, dev, count, dtype, (void *)ctx->stream);
    if (rc != 0) return set_err(ctx, LPB200_E_COMM, "all-reduce callback failed (%d)", rc);
    return 0;
}

// A failure on one shard (optim's non-finite error on a local gene, fitMeanDispersion.R:384-389) must reach every
// shard before any of them returns: the others would otherwise wait in the next collective forever.  The
// reference aborts the whole run on this error; so do all shards, together.
static int sync_error(lpb200_ctx *ctx, int rc) {
    if (ctx->world <= 1) return rc;
    double v = rc != 0 ? 1.0 : 0.0;
    if (cudaMemcpyAsync(ctx->d_xchg + 8, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return rc ? rc : LPB200_E_CUDA;
    int rc2 = do_allreduce(ctx, ctx->d_xchg + 8, 1, LPB200_F64);
    if (rc2) return rc ? rc : rc2;
    if (cudaMemcpyAsync(&v, ctx->d_xchg + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return rc ? rc : LPB200_E_CUDA;
    if (rc == 0 && v > 0) return set_err(ctx, LPB200_E_NONFINITE, "another gene shard reported an error (non-finite objective); the run is aborted on every shard");
    return rc;
}

// ---------------------------------------------------------------------------------------------
// data
// ---------------------------------------------------------------------------------------------
// temporary int32 image of the count matrix: freed on every error path, released into the context
// only when it becomes the resident copy
struct TmpCounts {
    int32_t *p = nullptr;
    ~TmpCounts() { if (p) cudaFree(p); }
    int32_t *release() { int32_t *q = p; p = nullptr; return q; }
};

static void free_design(lpb200_ctx *c);

static int finish_counts(lpb200_ctx *ctx, TmpCounts &tmp, int G, int N, size_t pitch) {
    int32_t *d_i32 = tmp.p;
    free_design(ctx);  // a design belongs to one matrix shape
    // drop the previous matrix first: a failure below leaves the context in the "counts not loaded" state
    if (ctx->d_counts) { cudaFree(ctx->d_counts); ctx->d_counts = nullptr; }
    if (ctx->d_xmax) { cudaFree(ctx->d_xmax); ctx->d_xmax = nullptr; }
    if (ctx->d_perm) { cudaFree(ctx->d_perm); ctx->d_perm = nullptr; }
    if (ctx->d_nnz) { cudaFree(ctx->d_nnz); ctx->d_nnz = nullptr; }
    if (ctx->d_nobs) { cudaFree(ctx->d_nobs); ctx->d_nobs = nullptr; }
    ctx->G = ctx->N = 0;
    // row statistics, then narrow to uint16 when every count fits
    DevBuf<double> d_mean;
    CK(d_mean.alloc(G));
    CK(cudaMalloc((void **)&ctx->d_xmax, sizeof(int32_t) * G));
    row_stats_kernel<<<G, 256, 0, ctx->stream>>>(d_i32, G, N, pitch, d_mean.p, ctx->d_xmax);
    CK(cudaGetLastError());
    CK(cudaMalloc((void **)&ctx->d_perm, sizeof(uint32_t) * (size_t)G * pitch));
    CK(cudaMalloc((void **)&ctx->d_nnz, sizeof(int32_t) * G));
    CK(cudaMalloc((void **)&ctx->d_nobs, sizeof(int32_t) * G));
    build_perm_kernel<<<G, 256, 0, ctx->stream>>>(d_i32, G, N, pitch, ctx->d_perm, ctx->d_nnz, ctx->d_nobs);
    CK(cudaGetLastError());
    DevBuf<int32_t> d_colmax;
    DevBuf<unsigned long long> d_neg;
    CK(d_colmax.alloc(N));
    CK(d_neg.alloc(1));
    CK(cudaMemsetAsync(d_neg.p, 0, sizeof(unsigned long long), ctx->stream));
    col_stats_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(d_i32, G, N, pitch, d_colmax.p, d_neg.p);
    CK(cudaGetLastError());
    ctx->row_means.resize(G);
    ctx->row_max.resize(G);
    ctx->col_max.resize(N);
    std::vector<int32_t> &xmax = ctx->row_max;
    CK(cudaMemcpyAsync(ctx->row_means.data(), d_mean.p, sizeof(double) * G, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(xmax.data(), ctx->d_xmax, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->col_max.data(), d_colmax.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&ctx->n_negative, d_neg.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->global_max = 0;
    for (int v : xmax) ctx->global_max = std::max(ctx->global_max, v);
    if (ctx->global_max < (int)LP_NA_U16) {
        uint16_t *d16 = nullptr;
        CK(cudaMalloc((void **)&d16, sizeof(uint16_t) * (size_t)G * pitch));
        narrow_u16_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_i32, d16, (size_t)G * pitch);
        cudaError_t e1 = cudaGetLastError();
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(ctx->stream);
        if (e1 != cudaSuccess) {
            cudaFree(d16);
            return set_err(ctx, LPB200_E_CUDA, "narrow_u16_kernel failed: %s", cudaGetErrorString(e1));
        }
        ctx->d_counts = d16;   // tmp frees the int32 image
        ctx->is_u16 = true;
    } else {
        ctx->d_counts = tmp.release();
        ctx->is_u16 = false;
    }
    ctx->G = G;
    ctx->N = N;
    ctx->pitch = pitch;
    ctx->stats.launches += 4;
    ctx->have_starts = false;
    return 0;
}

extern "C" int lpb200_count_flags(lpb200_ctx *ctx, int32_t *gene_max, int32_t *cell_max, int64_t *n_negative) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) {
        if (ctx->bounds.empty()) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
        if (n_negative) *n_negative = 0;
        if (cell_max) std::fill(cell_max, cell_max + ctx->N, 0);
        for (size_t k = 0; k < ctx->shards.size(); k++) {
            const lpb200_ctx *c = ctx->shards[k];
            if (gene_max) memcpy(gene_max + ctx->bounds[k], c->row_max.data(), sizeof(int32_t) * c->G);
            if (cell_max)
                for (int j = 0; j < ctx->N; j++) cell_max[j] = std::max(cell_max[j], c->col_max[j]);
            if (n_negative) *n_negative += (int64_t)c->n_negative;
        }
        return 0;
    }
    if (!ctx->d_counts) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
    if (gene_max) memcpy(gene_max, ctx->row_max.data(), sizeof(int32_t) * ctx->G);
    if (cell_max) memcpy(cell_max, ctx->col_max.data(), sizeof(int32_t) * ctx->N);
    if (n_negative) *n_negative = (int64_t)ctx->n_negative;
    return 0;
}

static size_t row_pitch(int N) { return ((size_t)N + 63) / 64 * 64; }

// rows [0, G) of an R matrix with leading dimension ld >= G (a gene shard of a larger matrix when ld > G)
static int load_counts_dense_ld(lpb200_ctx *ctx, int G, int N, const double *x, size_t ld) {
    if (!ctx) return LPB200_E_CUDA;
    if (G <= 0 || N <= 0 || !x) return set_err(ctx, LPB200_E_ARG, "load_counts_dense: bad arguments");
    CK(cudaSetDevice(ctx->device));
    size_t pitch = row_pitch(N);
    TmpCounts tmp;
    CK(cudaMalloc((void **)&tmp.p, sizeof(int32_t) * (size_t)G * pitch));
    int32_t *d_i32 = tmp.p;
    CK(cudaMemsetAsync(d_i32, 0, sizeof(int32_t) * (size_t)G * pitch, ctx->stream));
    // column slabs of the R matrix are contiguous: stream them through a bounded staging buffer
    size_t slab_cols = std::max<size_t>(1, std::min<size_t>(N, ((size_t)256 << 20) / (sizeof(double) * (size_t)G)));
    DevBuf<double> d_slab;
    CK(d_slab.alloc(slab_cols * G));
    for (size_t j0 = 0; j0 < (size_t)N; j0 += slab_cols) {
        size_t nj = std::min(slab_cols, (size_t)N - j0);
        if (ld == (size_t)G)
            CK(cudaMemcpyAsync(d_slab.p, x + j0 * G, sizeof(double) * nj * G, cudaMemcpyHostToDevice, ctx->stream));
        else
            CK(cudaMemcpy2DAsync(d_slab.p, sizeof(double) * G, x + j0 * ld, sizeof(double) * ld, sizeof(double) * G, nj,
                                 cudaMemcpyHostToDevice, ctx->stream));
        dim3 grid((G + 31) / 32, (unsigned)((nj + 31) / 32)), block(32, 8);
        pack_dense_kernel<<<grid, block, 0, ctx->stream>>>(d_slab.p, G, (int)j0, (int)nj, d_i32, pitch);
        CK(cudaGetLastError());
        ctx->stats.launches++;
    }
    return finish_counts(ctx, tmp, G, N, pitch);
}

extern "C" int lpb200_load_counts_dense(lpb200_ctx *ctx, int G, int N, const double *x) {
    if (is_multi(ctx)) return multi_load_counts_dense(ctx, G, N, x);
    return load_counts_dense_ld(ctx, G, N, x, (size_t)std::max(G, 0));
}

extern "C" int lpb200_load_counts_csc(lpb200_ctx *ctx, int G, int N, const int32_t *p, const int32_t *i, const double *x) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) return multi_load_counts_csc(ctx, G, N, p, i, x);
    if (G <= 0 || N <= 0 || !p) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: bad arguments");
    if (p[0] != 0 || p[N] < 0 || (p[N] > 0 && (!i || !x))) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: bad column pointers");
    for (int j = 0; j < N; j++)
        if (p[j + 1] < p[j]) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: column pointers decrease at column %d", j);
    for (int32_t e = 0; e < p[N]; e++)
        if (i[e] < 0 || i[e] >= G) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: row index %d out of range", (int)i[e]);
    CK(cudaSetDevice(ctx->device));
    size_t nnz = (size_t)p[N];
    size_t pitch = row_pitch(N);
    TmpCounts tmp;
    CK(cudaMalloc((void **)&tmp.p, sizeof(int32_t) * (size_t)G * pitch));
    int32_t *d_i32 = tmp.p;
    CK(cudaMemsetAsync(d_i32, 0, sizeof(int32_t) * (size_t)G * pitch, ctx->stream));
    DevBuf<int32_t> dp, di;
    DevBuf<double> dx;
    CK(dp.alloc(N + 1));
    CK(di.alloc(nnz));
    CK(dx.alloc(nnz));
    CK(cudaMemcpyAsync(dp.p, p, sizeof(int32_t) * (N + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        CK(cudaMemcpyAsync(di.p, i, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dx.p, x, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    scatter_csc_kernel<<<N, 128, 0, ctx->stream>>>(dp.p, di.p, dx.p, N, d_i32, pitch);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return finish_counts(ctx, tmp, G, N, pitch);
}

extern "C" int lpb200_load_counts_i32(lpb200_ctx *ctx, int G, int N, const int32_t *x, int is_device) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) return multi_load_counts_i32(ctx, G, N, x, is_device);
    if (G <= 0 || N <= 0 || !x) return set_err(ctx, LPB200_E_ARG, "load_counts_i32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    size_t pitch = row_pitch(N);
    TmpCounts tmp;
    CK(cudaMalloc((void **)&tmp.p, sizeof(int32_t) * (size_t)G * pitch));
    int32_t *d_i32 = tmp.p;
    CK(cudaMemsetAsync(d_i32, 0, sizeof(int32_t) * (size_t)G * pitch, ctx->stream));
    CK(cudaMemcpy2DAsync(d_i32, pitch * sizeof(int32_t), x, (size_t)N * sizeof(int32_t), (size_t)N * sizeof(int32_t), G,
                         is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return finish_counts(ctx, tmp, G, N, pitch);
}

extern "C" int lpb200_row_means(lpb200_ctx *ctx, double *means) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) return multi_row_means(ctx, means);
    if (!ctx->d_counts) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
    memcpy(means, ctx->row_means.data(), sizeof(double) * ctx->G);
    return 0;
}

// thin QR of an n x k column-major matrix by twice-iterated Gram-Schmidt (long double dots)
static bool thin_qr(const double *B, int n, int k, std::vector<double> &Q, std::vector<double> &R) {
    Q.assign(B, B + (size_t)n * k);
    R.assign((size_t)k * k, 0.0);
    for (int c = 0; c < k; c++) {
        double *q = Q.data() + (size_t)c * n;
        for (int pass = 0; pass < 2; pass++)
            for (int p = 0; p < c; p++) {
                const double *qp = Q.data() + (size_t)p * n;
                long double dot = 0.0L;
                for (int r = 0; r < n; r++) dot += (long double)qp[r] * q[r];
                for (int r = 0; r < n; r++) q[r] -= (double)dot * qp[r];
                R[p + (size_t)c * k] += (double)dot;
            }
        long double nrm = 0.0L;
        for (int r = 0; r < n; r++) nrm += (long double)q[r] * q[r];
        double nr = (double)sqrtl(nrm);
        if (!(nr > 0)) return false;
        R[c + (size_t)c * k] = nr;
        for (int r = 0; r < n; r++) q[r] /= nr;
    }
    return true;
}

template <typename T>
static int upload(lpb200_ctx *ctx, const T *h, size_t n, const T **d_out) {
    T *d = nullptr;
    CK(cudaMalloc((void **)&d, std::max<size_t>(n, 1) * sizeof(T)));
    ctx->design_allocs.push_back(d);
    CK(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = d;
    return 0;
}

extern "C" int lpb200_set_design(lpb200_ctx *ctx, const lpb200_design *d) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) return multi_set_design(ctx, d);
    if (!d) return set_err(ctx, LPB200_E_ARG, "design is NULL");
    if (!ctx->d_counts) return set_err(ctx, LPB200_E_STATE, "load counts before the design");
    if (d->n_genes != ctx->G || d->n_cells != ctx->N)
        return set_err(ctx, LPB200_E_ARG, "design is %d x %d but counts are %d x %d", d->n_genes, d->n_cells, ctx->G, ctx->N);
    if (d->n_conf_mu > LPB200_MAX_CONF || d->n_conf_disp > LPB200_MAX_CONF || d->n_pi_pred > LPB200_MAX_PI_PRED ||
        d->n_conf_mu < 0 || d->n_conf_disp < 0 || d->n_pi_pred < 0)
        return set_err(ctx, LPB200_E_ARG, "too many confounders / drop-out predictors");
    CK(cudaSetDevice(ctx->device));
    free_design(ctx);
    DevDesign &D = ctx->D;
    D = DevDesign{};
    const int G = ctx->G, N = ctx->N;
    D.G = G;
    D.N = N;
    int rc;
    if (d->size_factors) {
        bool all_one = true;
        for (int j = 0; j < N; j++)
            if (d->size_factors[j] != 1.0) { all_one = false; break; }
        if (!all_one && (rc = upload(ctx, d->size_factors, N, &D.sf))) return rc;
    }
    ctx->h_t.clear();
    if (d->pseudotime) {
        if ((rc = upload(ctx, d->pseudotime, N, &D.t))) return rc;
        ctx->h_t.assign(d->pseudotime, d->pseudotime + N);
    }
    if (d->group_idx) {
        for (int j = 0; j < N; j++)
            if (d->group_idx[j] < 0 || d->group_idx[j] >= d->n_groups) return set_err(ctx, LPB200_E_ARG, "group index out of range");
        if ((rc = upload(ctx, d->group_idx, N, &D.group))) return rc;
        D.n_groups = d->n_groups;
    }
    ctx->h_bs_mu.clear();
    if (d->spline_basis_mu && d->k_spline_mu > 0) {
        if ((rc = upload(ctx, d->spline_basis_mu, (size_t)N * d->k_spline_mu, &D.bs_mu))) return rc;
        D.k_mu = d->k_spline_mu;
        ctx->h_bs_mu.assign(d->spline_basis_mu, d->spline_basis_mu + (size_t)N * d->k_spline_mu);
    }
    if (d->spline_basis_disp && d->k_spline_disp > 0) {
        if ((rc = upload(ctx, d->spline_basis_disp, (size_t)N * d->k_spline_disp, &D.bs_disp))) return rc;
        D.k_disp = d->k_spline_disp;
    }
    D.n_conf_mu = d->n_conf_mu;
    if (d->n_conf_mu > 0) {
        for (int c = 0; c < d->n_conf_mu; c++) {
            D.nb_mu[c] = d->n_batches_mu[c];
            for (int j = 0; j < N; j++) {
                int b = d->batch_idx_mu[(size_t)c * N + j];
                if (b < 0 || b >= D.nb_mu[c]) return set_err(ctx, LPB200_E_ARG, "mu batch index out of range");
            }
        }
        if ((rc = upload(ctx, d->batch_idx_mu, (size_t)N * d->n_conf_mu, &D.bidx_mu))) return rc;
    }
    D.n_conf_disp = d->n_conf_disp;
    if (d->n_conf_disp > 0) {
        for (int c = 0; c < d->n_conf_disp; c++) {
            D.nb_disp[c] = d->n_batches_disp[c];
            for (int j = 0; j < N; j++) {
                int b = d->batch_idx_disp[(size_t)c * N + j];
                if (b < 0 || b >= D.nb_disp[c]) return set_err(ctx, LPB200_E_ARG, "dispersion batch index out of range");
            }
        }
        if ((rc = upload(ctx, d->batch_idx_disp, (size_t)N * d->n_conf_disp, &D.bidx_disp))) return rc;
    }
    if (d->pi_const_pred && d->n_pi_pred > 0) {
        if ((rc = upload(ctx, d->pi_const_pred, (size_t)G * d->n_pi_pred, &D.pi_pred))) return rc;
        D.P = d->n_pi_pred;
    }
    if (d->impulse_basis) {
        std::vector<double> Q, R;
        if (!thin_qr(d->impulse_basis, N, 4, Q, R)) return set_err(ctx, LPB200_E_ARG, "impulse_basis is rank deficient");
        const double *dq = nullptr;
        if ((rc = upload(ctx, Q.data(), (size_t)N * 4, &dq))) return rc;
        ctx->d_Q = const_cast<double *>(dq);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_design = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// model bookkeeping
// ---------------------------------------------------------------------------------------------
static int make_model(lpb200_ctx *ctx, const lpb200_model *m, DevModel &M) {
    const DevDesign &D = ctx->D;
    if (m->mu_model == LPB200_MU_MM)  // fitMeanDispersion.R:804, evalLogLik.R:271: stop("... not yet supported")
        return set_err(ctx, LPB200_E_NOTSUP, "LineagePulse ERROR: strMuModel=\"MM\" not yet supported (reference stub)");
    M = DevModel{};
    M.mu_model = m->mu_model;
    M.disp_model = m->disp_model;
    M.drop_model = m->drop_model;
    M.p_mu = lp_n_mu_params(m->mu_model, D.k_mu, D.n_groups);
    M.p_disp = lp_n_disp_params(m->disp_model, D.k_disp, D.n_groups);
    if (M.p_mu <= 0 || M.p_disp <= 0) return set_err(ctx, LPB200_E_ARG, "model needs covariates the design does not have");
    if (m->drop_model != LPB200_DROP_NONE && m->drop_model != LPB200_DROP_LOGISTIC && m->drop_model != LPB200_DROP_LOGISTIC_OFMU)
        return set_err(ctx, LPB200_E_ARG, "unknown drop-out model %d", m->drop_model);
    if (m->mu_model == LPB200_MU_IMPULSE && !D.t) return set_err(ctx, LPB200_E_ARG, "impulse model needs pseudotime");
    M.q_drop = m->drop_model == LPB200_DROP_NONE ? 0 : 1 + (m->drop_model == LPB200_DROP_LOGISTIC_OFMU ? 1 : 0) + D.P;
    for (int c = 0; c < D.n_conf_mu; c++) M.nbf_mu += D.nb_mu[c];
    for (int c = 0; c < D.n_conf_disp; c++) M.nbf_disp += D.nb_disp[c];
    M.n_theta = M.p_disp + (M.nbf_disp - D.n_conf_disp) + M.p_mu + (M.nbf_mu - D.n_conf_mu);
    if (M.n_theta > LPB200_MAX_PARAMS || M.p_mu > LP_MAX_PMU || M.p_disp > LP_MAX_PDISP || M.nbf_mu > LP_MAX_BMU ||
        M.nbf_disp > LP_MAX_BDISP || M.q_drop > LP_PI_NP)
        return set_err(ctx, LPB200_E_ARG, "model too large: n_theta=%d p_mu=%d p_disp=%d batch=%d/%d q=%d", M.n_theta, M.p_mu,
                       M.p_disp, M.nbf_mu, M.nbf_disp, M.q_drop);
    return 0;
}

// the parameter matrices a model reads or writes must be there (R would stop with "subscript out of bounds")
static int check_params(lpb200_ctx *ctx, const DevModel &M, const lpb200_params *p, bool need_drop) {
    if (!p || !p->mat_mu || !p->mat_disp) return set_err(ctx, LPB200_E_ARG, "mat_mu / mat_disp is NULL");
    if (ctx->D.n_conf_mu > 0 && !p->batch_mu) return set_err(ctx, LPB200_E_ARG, "batch_mu is NULL but the design has mean confounders");
    if (ctx->D.n_conf_disp > 0 && !p->batch_disp)
        return set_err(ctx, LPB200_E_ARG, "batch_disp is NULL but the design has dispersion confounders");
    if (need_drop && M.drop_model != LPB200_DROP_NONE && !p->mat_drop)
        return set_err(ctx, LPB200_E_ARG, "mat_drop is NULL but the drop-out model is not \"none\"");
    return 0;
}

static const lpb200_control *control_or_default(const lpb200_control *c, lpb200_control &tmp) {
    if (c) return c;
    lpb200_default_control(&tmp);
    return &tmp;
}

static int require_ready(lpb200_ctx *ctx) {
    if (!ctx) return LPB200_E_CUDA;
    if (!ctx->d_counts) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
    if (!ctx->have_design) return set_err(ctx, LPB200_E_STATE, "design not set");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return set_err(ctx, LPB200_E_CUDA, "cudaSetDevice failed");
    return 0;
}

// R matrices (column-major) <-> nat vectors [G][n_nat]
static void params_to_nat(const lpb200_ctx *ctx, const DevModel &M, const lpb200_params *p, std::vector<double> &nat) {
    const DevDesign &D = ctx->D;
    const int G = ctx->G, n_nat = lp_n_nat(M);
    nat.resize((size_t)G * n_nat);
    for (int i = 0; i < G; i++) {
        double *v = nat.data() + (size_t)i * n_nat;
        int u = 0;
        for (int k = 0; k < M.p_disp; k++) v[u++] = p->mat_disp[i + (size_t)k * G];
        size_t off = 0;
        for (int c = 0; c < D.n_conf_disp; c++) {
            for (int b = 0; b < D.nb_disp[c]; b++) v[u++] = p->batch_disp[off + i + (size_t)b * G];
            off += (size_t)G * D.nb_disp[c];
        }
        for (int k = 0; k < M.p_mu; k++) v[u++] = p->mat_mu[i + (size_t)k * G];
        off = 0;
        for (int c = 0; c < D.n_conf_mu; c++) {
            for (int b = 0; b < D.nb_mu[c]; b++) v[u++] = p->batch_mu[off + i + (size_t)b * G];
            off += (size_t)G * D.nb_mu[c];
        }
    }
}

static void nat_to_params(const lpb200_ctx *ctx, const DevModel &M, const double *v, int i, lpb200_params *p) {
    const DevDesign &D = ctx->D;
    const int G = ctx->G;
    int u = 0;
    for (int k = 0; k < M.p_disp; k++) p->mat_disp[i + (size_t)k * G] = v[u++];
    size_t off = 0;
    for (int c = 0; c < D.n_conf_disp; c++) {
        for (int b = 0; b < D.nb_disp[c]; b++) p->batch_disp[off + i + (size_t)b * G] = v[u++];
        off += (size_t)G * D.nb_disp[c];
    }
    for (int k = 0; k < M.p_mu; k++) p->mat_mu[i + (size_t)k * G] = v[u++];
    off = 0;
    for (int c = 0; c < D.n_conf_mu; c++) {
        for (int b = 0; b < D.nb_mu[c]; b++) p->batch_mu[off + i + (size_t)b * G] = v[u++];
        off += (size_t)G * D.nb_mu[c];
    }
}

static CountView count_view(const lpb200_ctx *ctx) {
    CountView C;
    C.ptr = ctx->d_counts;
    C.pitch = ctx->pitch;
    C.xmax = ctx->d_xmax;
    C.perm = ctx->d_perm;
    C.nnz = ctx->d_nnz;
    C.nobs = ctx->d_nobs;
    return C;
}

static int fit_threads(int N) { return N >= 2048 ? LP_FIT_THREADS : 128; }

// drop-out inputs shared by the fit / eval kernels
struct DropBufs {
    DevBuf<double> drop, pi, l1m, pi_scr, l1m_scr;
};
static int setup_drop(lpb200_ctx *ctx, const DevModel &M, const lpb200_params *p, int grid, FitArgs &A, DropBufs &B,
                      cudaStream_t stream) {
    const int N = ctx->N;
    A.drop = nullptr;
    A.pi_glob = A.l1m_glob = nullptr;
    A.pi_scratch = A.l1m_scratch = nullptr;
    if (M.drop_model == LPB200_DROP_NONE) return 0;
    if (!p->mat_drop) return set_err(ctx, LPB200_E_ARG, "mat_drop is NULL but the drop-out model is not \"none\"");
    CK(B.drop.alloc((size_t)N * M.q_drop));
    CK(cudaMemcpyAsync(B.drop.p, p->mat_drop, sizeof(double) * (size_t)N * M.q_drop, cudaMemcpyHostToDevice, stream));
    A.drop = B.drop.p;
    if (M.drop_model == LPB200_DROP_LOGISTIC) {
        if (ctx->D.P == 0) {
            CK(B.pi.alloc(N));
            CK(B.l1m.alloc(N));
            pi_vector_kernel<<<(N + 255) / 256, 256, 0, stream>>>(ctx->D, M, B.drop.p, B.pi.p, B.l1m.p);
            CK(cudaGetLastError());
            ctx->stats.launches++;
            A.pi_glob = B.pi.p;
            A.l1m_glob = B.l1m.p;
        } else {
            CK(B.pi_scr.alloc((size_t)grid * N));
            CK(B.l1m_scr.alloc((size_t)grid * N));
            A.pi_scratch = B.pi_scr.p;
            A.l1m_scratch = B.l1m_scr.p;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// evalLogLikMatrix
// ---------------------------------------------------------------------------------------------
static int eval_loglik_impl(lpb200_ctx *ctx, const DevModel &M, const lpb200_params *p, double *ll) {
    const int G = ctx->G, N = ctx->N;
    std::vector<double> nat;
    params_to_nat(ctx, M, p, nat);
    DevBuf<double> d_nat, d_ll;
    CK(d_nat.alloc(nat.size()));
    CK(d_ll.alloc(G));
    CK(cudaMemcpyAsync(d_nat.p, nat.data(), sizeof(double) * nat.size(), cudaMemcpyHostToDevice, ctx->stream));
    const int threads = fit_threads(N);
    const int grid = std::min(G, ctx->sm_count * 4);
    FitArgs A{};
    A.D = ctx->D;
    A.M = M;
    A.C = count_view(ctx);
    A.theta0 = d_nat.p;
    A.ll_out = d_ll.p;
    A.use_table = ctx->use_table;
    DropBufs B;
    int rc = setup_drop(ctx, M, p, grid, A, B, ctx->stream);
    if (rc) return rc;
    const size_t smem = sizeof(FitShared);
    if (ctx->is_u16) {
        CK(cudaFuncSetAttribute(eval_ll_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_ll_kernel<uint16_t><<<grid, threads, smem, ctx->stream>>>(A);
    } else {
        CK(cudaFuncSetAttribute(eval_ll_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_ll_kernel<int32_t><<<grid, threads, smem, ctx->stream>>>(A);
    }
    CK(cudaGetLastError());
    ctx->stats.launches++;
    ctx->stats.fn_evals += (uint64_t)G;
    ctx->stats.ole += (uint64_t)G * N;
    CK(cudaMemcpyAsync(ll, d_ll.p, sizeof(double) * G, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lpb200_eval_loglik(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, double *ll) {
    if (is_multi(ctx)) return multi_eval_loglik(ctx, m, p, ll);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m || !ll) return set_err(ctx, LPB200_E_ARG, "eval_loglik: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, p, true))) return rc;
    return eval_loglik_impl(ctx, M, p, ll);
}

// ---------------------------------------------------------------------------------------------
// initialiseImpulseParameters (cached: depends only on the data)
// ---------------------------------------------------------------------------------------------
static int ensure_starts(lpb200_ctx *ctx) {
    if (ctx->have_starts) return 0;
    if (!ctx->d_Q || !ctx->D.t) return set_err(ctx, LPB200_E_ARG, "impulse starts need pseudotime and impulse_basis");
    const int G = ctx->G, N = ctx->N;
    const int grid = std::min(G, ctx->sm_count * 4);
    DevBuf<double> d_peak, d_valley, d_fit;
    CK(d_peak.alloc((size_t)G * 7));
    CK(d_valley.alloc((size_t)G * 7));
    CK(d_fit.alloc((size_t)grid * N));
    StartsArgs A{};
    A.D = ctx->D;
    A.C = count_view(ctx);
    A.Q = ctx->d_Q;
    A.tmin = *std::min_element(ctx->h_t.begin(), ctx->h_t.end());
    A.tmax = *std::max_element(ctx->h_t.begin(), ctx->h_t.end());
    A.t_first = ctx->h_t.front();
    A.t_last = ctx->h_t.back();
    A.peak = d_peak.p;
    A.valley = d_valley.p;
    A.fit_scratch = d_fit.p;
    if (ctx->is_u16) impulse_starts_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(A);
    else impulse_starts_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    ctx->h_peak.resize((size_t)G * 7);
    ctx->h_valley.resize((size_t)G * 7);
    CK(cudaMemcpyAsync(ctx->h_peak.data(), d_peak.p, sizeof(double) * G * 7, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_valley.data(), d_valley.p, sizeof(double) * G * 7, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_starts = true;
    return 0;
}

extern "C" int lpb200_impulse_starts(lpb200_ctx *ctx, double *peak, double *valley) {
    if (is_multi(ctx)) return multi_impulse_starts(ctx, peak, valley);
    int rc = require_ready(ctx);
    if (rc) return rc;
    if ((rc = ensure_starts(ctx))) return rc;
    memcpy(peak, ctx->h_peak.data(), sizeof(double) * ctx->G * 7);
    memcpy(valley, ctx->h_valley.data(), sizeof(double) * ctx->G * 7);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// fitMuDisp
// ---------------------------------------------------------------------------------------------
// One fitMuDisp call = one job: launch (asynchronous, on the job's stream) and collect (blocks on that
// stream, selects the best start, writes the reference's outputs).  Independent jobs (the ZINB and NB
// fits of fitLPModels) are launched on separate streams so that the persistent CTAs of the next kernel
// fill the SMs the previous kernel's item-length tail leaves idle.
struct FitJob {
    DevModel M{};
    lpb200_params *p = nullptr;
    double *ll = nullptr;
    int32_t *conv = nullptr;
    int n = 0, n_starts = 1, grid = 0;
    size_t n_items = 0;
    std::vector<double> theta0;
    DevBuf<double> d_theta0, d_theta_out, d_ll;
    DevBuf<int> d_conv, d_queue;
    DevBuf<unsigned> d_evals;
    DropBufs B;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint64_t ole = 0, fn_evals = 0;
    // optional per-item outputs (lpb200_fit_mudisp_items): [n_items][n] / [n_items]
    double *items_theta0 = nullptr, *items_theta = nullptr, *items_ll = nullptr;
    int32_t *items_conv = nullptr;
    uint32_t *items_evals = nullptr;
    ~FitJob() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

static int fit_launch(lpb200_ctx *ctx, FitJob &J, const lpb200_control *c) {
    const DevDesign &D = ctx->D;
    const DevModel &M = J.M;
    const int G = ctx->G, N = ctx->N, n = M.n_theta, n_nat = lp_n_nat(M);
    const bool impulse = M.mu_model == LPB200_MU_IMPULSE;
    J.n = n;
    J.n_starts = impulse ? 3 : 1;
    int rc;
    if (impulse && (rc = ensure_starts(ctx))) return rc;
    std::vector<double> nat;
    params_to_nat(ctx, M, J.p, nat);
    // theta0 per (gene, start); impulse start order Peak, Valley, Prior (fitMeanDispersion.R:663)
    J.n_items = (size_t)G * J.n_starts;
    J.theta0.resize(J.n_items * n);
    for (int i = 0; i < G; i++) {
        PointNat g;
        lp_nat_load(M, nat.data() + (size_t)i * n_nat, g);
        if (impulse) {
            double coords[3][7];
            for (int k = 0; k < 7; k++) {
                coords[0][k] = ctx->h_peak[i + (size_t)k * G];
                coords[1][k] = ctx->h_valley[i + (size_t)k * G];
                coords[2][k] = g.mu[k];
            }
            for (int k = 2; k < 5; k++) coords[2][k] = log(coords[2][k]);  // :621-622
            for (int s = 0; s < 3; s++) lp_pack_theta(D, M, g, coords[s], J.theta0.data() + ((size_t)i * 3 + s) * n);
        } else {
            lp_pack_theta(D, M, g, nullptr, J.theta0.data() + (size_t)i * n);
        }
    }
    CK(J.d_theta0.alloc(J.theta0.size()));
    CK(J.d_theta_out.alloc(J.theta0.size()));
    CK(J.d_ll.alloc(J.n_items));
    CK(J.d_conv.alloc(J.n_items));
    CK(J.d_evals.alloc(J.n_items));
    CK(J.d_queue.alloc(1));
    CK(cudaEventCreate(&J.e0));
    CK(cudaEventCreate(&J.e1));
    CK(cudaMemcpyAsync(J.d_theta0.p, J.theta0.data(), sizeof(double) * J.theta0.size(), cudaMemcpyHostToDevice, J.stream));
    CK(cudaMemsetAsync(J.d_queue.p, 0, sizeof(int), J.stream));
    const int threads = fit_threads(N);
    // specialised kernels for the hot configurations (scalar dispersion, drop-out "none" / fixed "logistic": constant or
    // impulse mean without confounders -- C1-C3; groups or splines mean with mean batch factors -- C4 / C5), generic otherwise
    const bool fast_base = !ctx->no_fastpath && lp_disp_is_scalar(D, M) &&
                           (M.drop_model == LPB200_DROP_NONE || M.drop_model == LPB200_DROP_LOGISTIC);
    const bool fast_ci = fast_base && D.n_conf_mu == 0 && (M.mu_model == LPB200_MU_CONSTANT || M.mu_model == LPB200_MU_IMPULSE);
    const bool fast_cov = fast_base && !fast_ci && (M.mu_model == LPB200_MU_GROUPS || M.mu_model == LPB200_MU_SPLINES ||
                                                    M.mu_model == LPB200_MU_CONSTANT);   // constant: with mean batch factors
    const bool fast = fast_ci || fast_cov;
    const int zinb_bit = (M.drop_model == LPB200_DROP_LOGISTIC) ? 1 : 0;
    const int variant = fast_ci ? (M.mu_model == LPB200_MU_IMPULSE ? 3 : 1) + zinb_bit
                                : (fast_cov ? (M.mu_model == LPB200_MU_GROUPS ? 5 : (M.mu_model == LPB200_MU_SPLINES ? 7 : 9)) + zinb_bit : 0);
    // CTAs per work item.  A thread-block cluster working on one item divides the duration of an item -- and with it
    // the item-length tail of the dynamic queue -- by the cluster size; the result does not depend on it
    // (virtual-warp summation order, lp_kernels.cuh).  Measured (profiles/README.md): it pays when a launch has few
    // items per CTA slot (592 genes x 100 000 cells: -9 % with 2 CTAs per item); with 25 items per slot the
    // concurrent B / C / D launches already back-fill each other's tails (2 CTAs per item: +-0 %, 4: +10 %).
    // Short rows stay with one CTA per item: the cluster barrier per evaluation would show against a 20-40 us pass.
    int cluster = ctx->cluster;
    if (cluster == 0) {
        const size_t slots = (size_t)ctx->sm_count * ctx->ctas_per_sm;
        cluster = (N >= 32768 && J.n_items < 8 * slots) ? 2 : 1;
    }
    cluster = std::min(cluster, LP_MAX_CLUSTER);
    if (fast_cov) cluster = 1;   // the covariate-model kernels are built for one CTA per item only
    if (threads != LP_FIT_THREADS) cluster = 1;
    const bool cl = cluster > 1;
    const size_t smem = fit_smem_bytes(cl);
    // persistent CTAs: as many as stay resident (64 registers per thread; shared memory allows 3 per SM)
    const int smem_limit = (int)((227 * 1024) / (smem + 1024));
    const int per_sm = std::max(1, std::min(ctx->ctas_per_sm * (LP_FIT_THREADS / threads), smem_limit));
    FitKernelFn kernel = fit_kernel(ctx->is_u16, variant, cl);
    CK(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)std::min<size_t>(J.n_items, (size_t)ctx->sm_count * per_sm);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = J.stream;
    if (cl) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3(cluster);
        int resident = 0;   // clusters that can be co-resident (GPC boundaries count)
        CK(cudaOccupancyMaxActiveClusters(&resident, (const void *)kernel, &cfg));
        resident = std::min(resident, ctx->sm_count * per_sm / cluster);
        grid = cluster * (int)std::max<size_t>(1, std::min<size_t>(J.n_items, (size_t)std::max(resident, 1)));
    }
    cfg.gridDim = dim3(grid);
    J.grid = grid / cluster;
    FitArgs A{};
    A.D = D;
    A.M = M;
    A.C = count_view(ctx);
    A.n_items = (int)J.n_items;
    A.n_starts = J.n_starts;
    A.queue = J.d_queue.p;
    A.theta0 = J.d_theta0.p;
    A.maxit = c->maxit_mudisp;
    A.reltol = c->reltol_mudisp;
    A.ndeps = c->ndeps;
    A.theta_out = J.d_theta_out.p;
    A.ll_out = J.d_ll.p;
    A.conv_out = J.d_conv.p;
    A.evals_out = J.d_evals.p;
    A.use_table = ctx->use_table;
    if ((rc = setup_drop(ctx, M, J.p, grid, A, J.B, J.stream))) return rc;
    CK(cudaEventRecord(J.e0, J.stream));
    CK(cudaLaunchKernelEx(&cfg, kernel, A));
    if (getenv("LPB200_DEBUG"))
        fprintf(stderr, "[lpb200] fit_mudisp: %s kernel (%s counts), mu=%d drop=%d items=%zu grid=%d threads=%d cluster=%d smem=%zu\n",
                fast ? "specialised" : "generic", ctx->is_u16 ? "u16" : "i32", M.mu_model, M.drop_model, J.n_items, grid,
                threads, cluster, smem);
    CK(cudaGetLastError());
    CK(cudaEventRecord(J.e1, J.stream));
    ctx->stats.launches++;
    return 0;
}

static int fit_collect(lpb200_ctx *ctx, FitJob &J, float *kernel_ms) {
    const DevDesign &D = ctx->D;
    const DevModel &M = J.M;
    const int G = ctx->G, N = ctx->N, n = J.n, n_starts = J.n_starts;
    const size_t n_items = J.n_items;
    std::vector<double> theta_out(J.theta0.size()), ll_items(n_items);
    std::vector<int> conv_items(n_items);
    std::vector<unsigned> evals(n_items);
    CK(cudaMemcpyAsync(theta_out.data(), J.d_theta_out.p, sizeof(double) * theta_out.size(), cudaMemcpyDeviceToHost, J.stream));
    CK(cudaMemcpyAsync(ll_items.data(), J.d_ll.p, sizeof(double) * n_items, cudaMemcpyDeviceToHost, J.stream));
    CK(cudaMemcpyAsync(conv_items.data(), J.d_conv.p, sizeof(int) * n_items, cudaMemcpyDeviceToHost, J.stream));
    CK(cudaMemcpyAsync(evals.data(), J.d_evals.p, sizeof(unsigned) * n_items, cudaMemcpyDeviceToHost, J.stream));
    CK(cudaStreamSynchronize(J.stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, J.e0, J.e1);
    if (kernel_ms) *kernel_ms = ms;
    if (getenv("LPB200_DEBUG")) {
        // how well the dynamic queue balanced this launch: greedy list schedule of the measured item
        // costs (objective evaluations) on `grid` CTA slots, ideal / makespan
        std::vector<double> slot(J.grid, 0.0);
        double total = 0.0;
        unsigned longest = 0;
        for (size_t it = 0; it < n_items; it++) {
            size_t k = std::min_element(slot.begin(), slot.end()) - slot.begin();
            slot[k] += evals[it];
            total += evals[it];
            longest = std::max(longest, evals[it]);
        }
        double makespan = *std::max_element(slot.begin(), slot.end());
        fprintf(stderr, "[lpb200] fit_mudisp: %zu items, %.0f evaluations/item (max %u), kernel %.1f ms, queue balance %.3f\n",
                n_items, total / n_items, longest, ms, total / J.grid / makespan);
    }
    if (J.items_theta0) memcpy(J.items_theta0, J.theta0.data(), sizeof(double) * J.theta0.size());
    if (J.items_theta) memcpy(J.items_theta, theta_out.data(), sizeof(double) * theta_out.size());
    if (J.items_ll) memcpy(J.items_ll, ll_items.data(), sizeof(double) * n_items);
    if (J.items_conv) memcpy(J.items_conv, conv_items.data(), sizeof(int32_t) * n_items);
    if (J.items_evals) memcpy(J.items_evals, evals.data(), sizeof(uint32_t) * n_items);
    bool nonfinite = false;
    for (int i = 0; i < G; i++) {
        int best = 0;
        bool any_bad = false;
        for (int s = 0; s < n_starts; s++) {
            size_t it = (size_t)i * n_starts + s;
            J.fn_evals += evals[it];
            J.ole += (uint64_t)evals[it] * N;
            if (conv_items[it] == 1001 || std::isnan(ll_items[it])) any_bad = true;
        }
        if (any_bad) {  // optim() error => the reference aborts (fitMeanDispersion.R:384-389)
            J.ll[i] = NAN;
            J.conv[i] = 1001;
            nonfinite = true;
            continue;
        }
        for (int s = 1; s < n_starts; s++)  // first maximum in the order Peak, Valley, Prior (:664-671)
            if (ll_items[(size_t)i * n_starts + s] > ll_items[(size_t)i * n_starts + best]) best = s;
        size_t it = (size_t)i * n_starts + best;
        PointNat g;
        lp_unpack_theta(D, M, theta_out.data() + it * n, g);  // same clamps as :391-498
        double v[LPB200_MAX_PARAMS * 3];
        lp_nat_store(M, g, v);
        nat_to_params(ctx, M, v, i, J.p);
        J.ll[i] = ll_items[it];
        J.conv[i] = conv_items[it];
    }
    ctx->stats.fn_evals += J.fn_evals;
    ctx->stats.ole += J.ole;
    if (nonfinite) return set_err(ctx, LPB200_E_NONFINITE, "fitMuDisp: non-finite objective (optim would have raised an error)");
    return 0;
}

static int fit_mudisp_impl(lpb200_ctx *ctx, const DevModel &M, lpb200_params *p, const lpb200_control *c, double *ll,
                           int32_t *conv, const FitJob *items = nullptr) {
    FitJob J;
    J.M = M;
    J.p = p;
    J.ll = ll;
    J.conv = conv;
    J.stream = ctx->stream;
    if (items) {
        J.items_theta0 = items->items_theta0;
        J.items_theta = items->items_theta;
        J.items_ll = items->items_ll;
        J.items_conv = items->items_conv;
        J.items_evals = items->items_evals;
    }
    int rc = fit_launch(ctx, J, c);
    if (rc) return rc;
    float ms = 0.f;
    rc = fit_collect(ctx, J, &ms);
    ctx->stats.kernel_ms += ms;
    return rc;
}

extern "C" int lpb200_fit_mudisp(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll,
                                 int32_t *conv) {
    if (is_multi(ctx)) return multi_fit_mudisp_items(ctx, m, p, c, ll, conv, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_mudisp: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, p, true))) return rc;
    lpb200_control tmp;
    return fit_mudisp_impl(ctx, M, p, control_or_default(c, tmp), ll, conv);
}

extern "C" int lpb200_fit_mudisp_items(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c,
                                       double *ll, int32_t *conv, int32_t *n_starts, int32_t *n_theta, double *theta0,
                                       double *theta, double *ll_items, int32_t *conv_items, uint32_t *evals_items) {
    if (is_multi(ctx)) return multi_fit_mudisp_items(ctx, m, p, c, ll, conv, n_starts, n_theta, theta0, theta, ll_items, conv_items, evals_items);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m || !ll || !conv || !n_starts || !n_theta) return set_err(ctx, LPB200_E_ARG, "fit_mudisp_items: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, p, true))) return rc;
    *n_starts = (M.mu_model == LPB200_MU_IMPULSE) ? 3 : 1;
    *n_theta = M.n_theta;
    FitJob items;
    items.items_theta0 = theta0;
    items.items_theta = theta;
    items.items_ll = ll_items;
    items.items_conv = conv_items;
    items.items_evals = evals_items;
    lpb200_control tmp;
    return fit_mudisp_impl(ctx, M, p, control_or_default(c, tmp), ll, conv, &items);
}

// ---------------------------------------------------------------------------------------------
// fitPi
// ---------------------------------------------------------------------------------------------
template <typename CT>
static void launch_pi_eval(lpb200_ctx *ctx, const PiArgs &A) {
    dim3 grid((ctx->N + 127) / 128, A.n_chunks);
    const bool hoist = A.M.drop_model == LPB200_DROP_LOGISTIC && A.D.P == 0;   // gene-independent drop-out rates
    const int kp = std::max(A.kp, A.shared_nk);
    if (hoist) pi_eval_kernel<CT, true, 2><<<grid, 128, 0, ctx->stream>>>(A);      // q = 1: at most 2 points per cell
    else if (kp <= 4) pi_eval_kernel<CT, false, 4><<<grid, 128, 0, ctx->stream>>>(A);
    else pi_eval_kernel<CT, false, LP_PI_KP><<<grid, 128, 0, ctx->stream>>>(A);
}

// Tail of the lock-step loop on a SHARDED run.  A cell that needs thousands of BFGS iterations would cost one
// all-reduce round trip per objective evaluation; instead the few remaining cells' count columns, the gene
// parameters and the gene-wise drop-out predictors of every shard are gathered once on every shard (a summing
// all-reduce of buffers in which each shard fills only its own gene range -- any transport that can all-reduce
// can do it), and every shard finishes all remaining cells with pi_tail_kernel over the full gene range.  The
// 16-gene summation blocks and their exact integer combination are those of the lock-step path, so the states
// stay identical on every shard and equal to a single-device run, bit for bit.
static int fit_pi_sharded_tail(lpb200_ctx *ctx, const PiArgs &A, const double *d_nat, const int *d_list, int n_tail) {
    const int G = ctx->G, world = ctx->world, n_nat = lp_n_nat(A.M), P = A.D.P;
    int rc;
    // gene ranges of the shards
    std::vector<long long> gs(world, 0);
    gs[ctx->rank] = G;
    DevBuf<long long> d_gs;
    CK(d_gs.alloc(world));
    CK(cudaMemcpyAsync(d_gs.p, gs.data(), sizeof(long long) * world, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = do_allreduce(ctx, d_gs.p, world, LPB200_I64))) return rc;
    CK(cudaMemcpyAsync(gs.data(), d_gs.p, sizeof(long long) * world, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    long long g_off = 0, g_tot = 0;
    for (int r = 0; r < world; r++) {
        if (r < ctx->rank) g_off += gs[r];
        g_tot += gs[r];
    }
    if (g_off % LP_PI_BLOCK != 0) return set_err(ctx, LPB200_E_ARG, "gene shards must start on multiples of %d genes", LP_PI_BLOCK);
    // gene parameters of all shards
    DevBuf<double> d_nat_all, d_pred_all;
    CK(d_nat_all.alloc((size_t)g_tot * n_nat));
    CK(cudaMemsetAsync(d_nat_all.p, 0, sizeof(double) * (size_t)g_tot * n_nat, ctx->stream));
    CK(cudaMemcpyAsync(d_nat_all.p + (size_t)g_off * n_nat, d_nat, sizeof(double) * (size_t)G * n_nat, cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = do_allreduce(ctx, d_nat_all.p, (size_t)g_tot * n_nat, LPB200_F64))) return rc;
    if (P > 0) {   // matPiConstPredictors: G x P column-major on every shard -> g_tot x P
        CK(d_pred_all.alloc((size_t)g_tot * P));
        CK(cudaMemsetAsync(d_pred_all.p, 0, sizeof(double) * (size_t)g_tot * P, ctx->stream));
        CK(cudaMemcpy2DAsync(d_pred_all.p + g_off, sizeof(double) * g_tot, A.D.pi_pred, sizeof(double) * G, sizeof(double) * G, P,
                             cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = do_allreduce(ctx, d_pred_all.p, (size_t)g_tot * P, LPB200_F64))) return rc;
    }
    // count columns of the remaining cells over all genes, cell-major (row length even: the buffer travels as int64 words)
    const size_t ld = ((size_t)g_tot + 1) / 2 * 2;
    DevBuf<int32_t> d_cols;
    CK(d_cols.alloc((size_t)n_tail * ld));
    CK(cudaMemsetAsync(d_cols.p, 0, sizeof(int32_t) * (size_t)n_tail * ld, ctx->stream));
    dim3 grid((unsigned)std::min(64, (G + 255) / 256), (unsigned)n_tail);
    if (ctx->is_u16) pi_gather_counts_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(A.C, G, d_list, n_tail, (int)g_off, d_cols.p, ld);
    else pi_gather_counts_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(A.C, G, d_list, n_tail, (int)g_off, d_cols.p, ld);
    CK(cudaGetLastError());
    if ((rc = do_allreduce(ctx, d_cols.p, (size_t)n_tail * ld / 2, LPB200_I64))) return rc;
    PiArgs T = A;
    T.D.G = (int)g_tot;
    T.D.pi_pred = P > 0 ? d_pred_all.p : nullptr;
    T.nat = d_nat_all.p;
    T.tail_counts = d_cols.p;
    T.tail_ld = ld;
    T.xphi_tab = nullptr;   // the per-gene (x, phi) tables cover this shard's genes only
    T.xphi_w = 0;
    if (ctx->is_u16) pi_tail_kernel<uint16_t><<<n_tail, 256, 0, ctx->stream>>>(T, d_list);
    else pi_tail_kernel<int32_t><<<n_tail, 256, 0, ctx->stream>>>(T, d_list);
    CK(cudaGetLastError());
    ctx->stats.launches += 2;
    CK(cudaStreamSynchronize(ctx->stream));   // the gathered buffers are released on return
    return 0;
}

static int fit_pi_impl(lpb200_ctx *ctx, const DevModel &M, int drop_fit_group, lpb200_params *p, const lpb200_control *c,
                       double *ll, int32_t *conv) {
    const int G = ctx->G, N = ctx->N, q = M.q_drop;
    if (M.drop_model == LPB200_DROP_NONE) return set_err(ctx, LPB200_E_ARG, "fitPi needs a drop-out model");
    if (!p->mat_drop) return set_err(ctx, LPB200_E_ARG, "mat_drop is NULL");
    std::vector<double> nat;
    params_to_nat(ctx, M, p, nat);
    DevBuf<double> d_nat, d_drop, d_vals, d_ll, d_total;
    DevBuf<long long> d_partial, d_vals_i;
    DevBuf<PiState> d_states;
    DevBuf<int> d_active, d_conv;
    DevBuf<unsigned> d_evals;
    const int cell_blocks = (N + 127) / 128;
    // gene chunks: enough CTAs for every SM to hold its 7 resident 128-thread CTAs of the 72-register instantiation
    int n_chunks = std::max(1, std::min(64, (ctx->sm_count * 8 + cell_blocks - 1) / cell_blocks));
    n_chunks = std::min(n_chunks, G);
    int gpc = (G + n_chunks - 1) / n_chunks;
    gpc = (gpc + LP_PI_BLOCK - 1) / LP_PI_BLOCK * LP_PI_BLOCK;   // chunks start on summation-block boundaries
    n_chunks = (G + gpc - 1) / gpc;
    CK(d_nat.alloc(nat.size()));
    CK(d_drop.alloc((size_t)N * q));
    const int kp = std::max(1, 2 * q);   // live point columns of a round (a gradient's 2q central-difference points)
    CK(d_partial.alloc((size_t)3 * n_chunks * kp * N));
    CK(d_vals_i.alloc((size_t)3 * kp * N));
    CK(d_vals.alloc((size_t)kp * N));
    CK(d_states.alloc(N));
    CK(d_active.alloc(1));
    CK(d_total.alloc(LP_PI_KP));
    CK(cudaMemcpyAsync(d_nat.p, nat.data(), sizeof(double) * nat.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_drop.p, p->mat_drop, sizeof(double) * (size_t)N * q, cudaMemcpyHostToDevice, ctx->stream));
    PiArgs A{};
    A.D = ctx->D;
    A.M = M;
    A.C = count_view(ctx);
    A.nat = d_nat.p;
    A.states = d_states.p;
    A.partial_i = d_partial.p;
    A.vals_i = d_vals_i.p;
    A.vals = d_vals.p;
    A.kp = kp;
    A.tail_counts = nullptr;
    A.tail_ld = 0;
    // (x, phi) part of dnbinom per gene: the gene parameters are fixed for the whole call
    DevBuf<double> d_xphi;
    A.xphi_tab = nullptr;
    A.xphi_w = 0;
    if (ctx->use_table && lp_disp_is_scalar(ctx->D, M)) {
        const int w = std::min(ctx->global_max + 1, 1024);
        if (w > 1) {
            CK(d_xphi.alloc((size_t)G * w));
            pi_xphi_table_kernel<<<G, 256, 0, ctx->stream>>>(d_nat.p, lp_n_nat(M), ctx->d_xmax, w, d_xphi.p);
            CK(cudaGetLastError());
            ctx->stats.launches++;
            A.xphi_tab = d_xphi.p;
            A.xphi_w = w;
        }
    }
    A.genes_per_chunk = gpc;
    A.n_chunks = n_chunks;
    A.shared_nk = 0;
    A.n_active = d_active.p;
    A.ndeps = c->ndeps;
    const int nb = (N + 127) / 128;
    int rc;
    PhaseTimer pt("fit_pi");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));

    if (drop_fit_group == LPB200_DROPFIT_PERCELL) {
        pi_start_kernel<<<nb, 128, 0, ctx->stream>>>(A, d_drop.p, c->maxit_pi, c->reltol_pi);
        CK(cudaGetLastError());
        ctx->stats.launches++;
        // Lock-step rounds.  The active-cell count is read back (a host sync) every round at first
        // and every 4th round once most cells have stopped; on one GPU the last few cells are handed
        // to pi_tail_kernel, which finishes their BFGS without any further host round trips.
        const int tail_threshold = 256;
        int since_check = 0, check_every = 1;
        DevBuf<int> d_list;
        CK(d_list.alloc(N));
        for (long round = 0;; round++) {
            CK(cudaMemsetAsync(d_active.p, 0, sizeof(int), ctx->stream));
            if (ctx->is_u16) launch_pi_eval<uint16_t>(ctx, A);
            else launch_pi_eval<int32_t>(ctx, A);
            CK(cudaGetLastError());
            pi_reduce_kernel<<<nb, 128, 0, ctx->stream>>>(A);
            CK(cudaGetLastError());
            if ((rc = do_allreduce(ctx, d_vals_i.p, (size_t)3 * kp * N, LPB200_I64))) return rc;   // the live columns only
            pi_combine_kernel<<<nb, 128, 0, ctx->stream>>>(A);
            CK(cudaGetLastError());
            pi_advance_kernel<<<nb, 128, 0, ctx->stream>>>(A);
            CK(cudaGetLastError());
            ctx->stats.launches += 4;
            if (++since_check < check_every) continue;
            since_check = 0;
            int active = 0;
            CK(cudaMemcpyAsync(&active, d_active.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (pt.on) {
                char buf[64];
                snprintf(buf, sizeof buf, "round %ld (%d cells active)", round, active);
                pt.lap(buf);
            }
            if (active == 0) break;
            if (round >= 8) check_every = 4;
            if (active <= tail_threshold) {
                pi_compact_kernel<<<1, 256, 0, ctx->stream>>>(A, d_list.p, d_active.p);
                CK(cudaGetLastError());
                ctx->stats.launches++;
                if (ctx->world == 1) {
                    if (ctx->is_u16) pi_tail_kernel<uint16_t><<<active, 256, 0, ctx->stream>>>(A, d_list.p);
                    else pi_tail_kernel<int32_t><<<active, 256, 0, ctx->stream>>>(A, d_list.p);
                    CK(cudaGetLastError());
                    ctx->stats.launches++;
                } else if ((rc = fit_pi_sharded_tail(ctx, A, d_nat.p, d_list.p, active))) {
                    return rc;
                }
                break;
            }
        }
        pt.lap("tail");
        CK(d_ll.alloc(N));
        CK(d_conv.alloc(N));
        CK(d_evals.alloc(N));
        pi_finish_kernel<<<nb, 128, 0, ctx->stream>>>(A, d_drop.p, d_ll.p, d_conv.p, d_evals.p);
        CK(cudaGetLastError());
        ctx->stats.launches++;
        std::vector<unsigned> evals(N);
        CK(cudaMemcpyAsync(p->mat_drop, d_drop.p, sizeof(double) * (size_t)N * q, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ll, d_ll.p, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(conv, d_conv.p, sizeof(int) * N, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(evals.data(), d_evals.p, sizeof(unsigned) * N, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        pt.lap("results");
        bool nonfinite = false;
        for (int j = 0; j < N; j++) {
            ctx->stats.fn_evals += evals[j];
            ctx->stats.ole += (uint64_t)evals[j] * G;
            if (conv[j] == 1001) nonfinite = true;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->stats.kernel_ms += ms;
        if (nonfinite) return set_err(ctx, LPB200_E_NONFINITE, "fitPi: non-finite objective (optim would have raised an error)");
        return 0;
    }

    // ForAllCells (fitDropout.R:480-499): one theta, objective = sum over the whole matrix
    PiState st;
    double th0[LP_PI_NP];
    for (int k = 0; k < q; k++) th0[k] = p->mat_drop[0 + (size_t)k * N];  // all rows are the same
    if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) th0[1] = log(-th0[1]);
    bfgs_start(st, q, th0, c->maxit_pi, c->reltol_pi, c->ndeps);
    double vals[LP_PI_KP];
    while (st.req != LP_REQ_NONE) {
        int nk = (st.req == LP_REQ_F) ? 1 : 2 * q;
        for (int k = 0; k < nk; k++) {
            for (int i = 0; i < q; i++) A.shared_theta[k * LP_PI_NP + i] = st.b[i];
            if (st.req == LP_REQ_GRAD) {
                int cc = bfgs_grad_coord(k);
                A.shared_theta[k * LP_PI_NP + cc] = bfgs_grad_point(st.b[cc], k, c->ndeps);
            }
        }
        A.shared_nk = nk;
        if (ctx->is_u16) launch_pi_eval<uint16_t>(ctx, A);
        else launch_pi_eval<int32_t>(ctx, A);
        CK(cudaGetLastError());
        pi_reduce_kernel<<<nb, 128, 0, ctx->stream>>>(A);
        CK(cudaGetLastError());
        if ((rc = do_allreduce(ctx, d_vals_i.p, (size_t)3 * kp * N, LPB200_I64))) return rc;
        pi_combine_kernel<<<nb, 128, 0, ctx->stream>>>(A);
        CK(cudaGetLastError());
        pi_total_kernel<<<nk, 256, 0, ctx->stream>>>(d_vals.p, N, d_total.p);   // every shard sums the same cells in the same order
        CK(cudaGetLastError());
        ctx->stats.launches += 4;
        CK(cudaMemcpyAsync(vals, d_total.p, sizeof(double) * nk, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->stats.fn_evals += (uint64_t)nk;
        ctx->stats.ole += (uint64_t)nk * G * N;
        bfgs_advance(st, vals);
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.kernel_ms += ms;
    conv[0] = st.fail;
    if (st.fail == 1001) {
        ll[0] = NAN;
        return set_err(ctx, LPB200_E_NONFINITE, "fitPi (ForAllCells): non-finite objective");
    }
    double lin[LP_PI_NP];
    lp_pi_transform(M.drop_model, q, st.b, lin);
    for (int k = 0; k < q; k++)
        for (int j = 0; j < N; j++) p->mat_drop[j + (size_t)k * N] = lin[k];
    ll[0] = -st.Fmin;
    return 0;
}

extern "C" int lpb200_fit_pi(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll,
                             int32_t *conv) {
    if (is_multi(ctx)) return multi_fit_pi(ctx, m, p, c, ll, conv);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_pi: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, p, true))) return rc;
    lpb200_control tmp;
    return fit_pi_impl(ctx, M, m->drop_fit_group, p, control_or_default(c, tmp), ll, conv);
}

// ---------------------------------------------------------------------------------------------
// initial values (initialiseMuModel.R:93-133, initialiseDispModel.R:89-130, initialiseDropModel.R:73-91)
// ---------------------------------------------------------------------------------------------
static int init_params_impl(lpb200_ctx *ctx, const DevModel &M, lpb200_params *out, double min_rowmean_global, bool init_drop) {
    const DevDesign &D = ctx->D;
    const int G = ctx->G, N = ctx->N;
    double minmean = INFINITY;
    for (int i = 0; i < G; i++) minmean = std::min(minmean, ctx->row_means[i]);
    if (!std::isnan(min_rowmean_global)) minmean = min_rowmean_global;
    for (int i = 0; i < G; i++) {
        double mi = ctx->row_means[i];
        if (mi < 1e-10) mi = 1e-10;
        if (M.mu_model == LPB200_MU_CONSTANT) out->mat_mu[i] = mi;
        else if (M.mu_model == LPB200_MU_IMPULSE) {
            for (int k = 0; k < 7; k++) out->mat_mu[i + (size_t)k * G] = (k >= 2 && k < 5) ? mi : 1.0;
        } else if (M.mu_model == LPB200_MU_GROUPS) {
            for (int k = 0; k < M.p_mu; k++) out->mat_mu[i + (size_t)k * G] = mi;
        }
        for (int k = 0; k < M.p_disp; k++)
            out->mat_disp[i + (size_t)k * G] = (M.disp_model == LPB200_DISP_SPLINES && k > 0) ? 0.0 : 1.0;
    }
    if (M.mu_model == LPB200_MU_SPLINES) {  // lm(log(x+1) ~ 0 + basis)$coef = R^-1 Q^T y
        const int K = M.p_mu;
        std::vector<double> Q, R;
        if (!thin_qr(ctx->h_bs_mu.data(), N, K, Q, R)) return set_err(ctx, LPB200_E_ARG, "spline basis is rank deficient");
        std::vector<double> Pinv((size_t)K * N);  // row-major K x N
        for (int j = 0; j < N; j++) {
            double col[LP_MAX_PMU];
            for (int cidx = K - 1; cidx >= 0; cidx--) {
                double s = Q[j + (size_t)cidx * N];
                for (int pp = cidx + 1; pp < K; pp++) s -= R[cidx + (size_t)pp * K] * col[pp];
                col[cidx] = s / R[cidx + (size_t)cidx * K];
            }
            for (int k = 0; k < K; k++) Pinv[(size_t)k * N + j] = col[k];
        }
        DevBuf<double> d_pinv, d_coef;
        CK(d_pinv.alloc(Pinv.size()));
        CK(d_coef.alloc((size_t)G * K));
        CK(cudaMemcpyAsync(d_pinv.p, Pinv.data(), sizeof(double) * Pinv.size(), cudaMemcpyHostToDevice, ctx->stream));
        const int grid = std::min(G, ctx->sm_count * 8);
        if (ctx->is_u16) spline_init_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(D, count_view(ctx), d_pinv.p, K, d_coef.p);
        else spline_init_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(D, count_view(ctx), d_pinv.p, K, d_coef.p);
        CK(cudaGetLastError());
        ctx->stats.launches++;
        CK(cudaMemcpyAsync(out->mat_mu, d_coef.p, sizeof(double) * (size_t)G * K, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (D.n_conf_mu > 0) {
        if (!out->batch_mu) return set_err(ctx, LPB200_E_ARG, "batch_mu is NULL");
        for (size_t k = 0; k < (size_t)G * M.nbf_mu; k++) out->batch_mu[k] = 1.0;
    }
    if (D.n_conf_disp > 0) {
        if (!out->batch_disp) return set_err(ctx, LPB200_E_ARG, "batch_disp is NULL");
        for (size_t k = 0; k < (size_t)G * M.nbf_disp; k++) out->batch_disp[k] = 1.0;
    }
    if (init_drop && M.drop_model != LPB200_DROP_NONE && out->mat_drop) {
        const double target = 0.99;
        for (int k = 0; k < M.q_drop; k++)
            for (int j = 0; j < N; j++) out->mat_drop[j + (size_t)k * N] = 0.0;
        if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) {
            const double slope = -1.0;
            const double off = log(target) - log(1 - target) - slope * log(minmean);
            for (int j = 0; j < N; j++) {
                out->mat_drop[j] = off;
                out->mat_drop[j + (size_t)N] = slope;
            }
        } else {
            const double off = log(target) - log(1 - target);
            for (int j = 0; j < N; j++) out->mat_drop[j] = off;
        }
    }
    return 0;
}

extern "C" int lpb200_init_params(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *out, double min_rowmean_global) {
    if (is_multi(ctx)) return multi_init_params(ctx, m, out, min_rowmean_global);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m) return set_err(ctx, LPB200_E_ARG, "init_params: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, out, false))) return rc;
    return init_params_impl(ctx, M, out, min_rowmean_global, true);
}

// ---------------------------------------------------------------------------------------------
// fitModel: the EM / coordinate-ascent loop (fitModel.R:137-356)
// ---------------------------------------------------------------------------------------------
// sum(vecLL) over all genes of all shards (fitModel.R:328): exact fixed-point accumulation, so every
// shard count gives the same scalar and the EM stopping rule takes the same decision
static int sum_ll(lpb200_ctx *ctx, const double *ll, int G, double *out) {
    ExactSum s = {0, 0, 0};
    for (int i = 0; i < G; i++) lp_exact_add(s, ll[i]);
    long long w[3] = {s.a, s.b, s.bad};
    if (ctx->world > 1) {
        CK(cudaMemcpyAsync(ctx->d_xchg, w, sizeof(w), cudaMemcpyHostToDevice, ctx->stream));
        int rc = do_allreduce(ctx, ctx->d_xchg, 3, LPB200_I64);
        if (rc) return rc;
        CK(cudaMemcpyAsync(w, ctx->d_xchg, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *out = lp_exact_value(w[0], w[1], w[2]);
    return 0;
}

extern "C" int lpb200_fit_model(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, int fit_drop,
                                const lpb200_control *c, double min_rowmean_global, double *em_ll, int32_t *n_iter,
                                int32_t *converged, double *ll_init, double *ll, int32_t *conv) {
    return lpb200_fit_model_from(ctx, m, p, fit_drop, c, min_rowmean_global, 0, em_ll, n_iter, converged, ll_init, ll, conv);
}

// fitModel with the reference's warm-start hooks (fitModel.R:147-151; initialiseMuModel.R:93,131-133,
// initialiseDispModel.R:89,109-129): the matrices named in keep_mask are the caller's initial values
// (matMuModelInit, lsmatBatchModelInitMu, matDispModelInit, lsmatBatchModelInitDisp); the others are initialised
// as initMuModel / initDispModel do.
extern "C" int lpb200_fit_model_from(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, int fit_drop,
                                     const lpb200_control *c, double min_rowmean_global, int keep_mask, double *em_ll,
                                     int32_t *n_iter, int32_t *converged, double *ll_init, double *ll, int32_t *conv) {
    if (is_multi(ctx)) return multi_fit_model(ctx, m, p, fit_drop, c, min_rowmean_global, keep_mask, em_ll, n_iter, converged, ll_init, ll, conv);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if (!m || !em_ll || !n_iter || !converged || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_model: NULL argument");
    if ((rc = make_model(ctx, m, M))) return rc;
    if ((rc = check_params(ctx, M, p, true))) return rc;
    lpb200_control ctl_default;
    c = control_or_default(c, ctl_default);
    const int G = ctx->G, N = ctx->N;
    if (fit_drop && M.drop_model == LPB200_DROP_NONE) fit_drop = 0;
    const int max_cycles = fit_drop ? c->max_cycles : 1;  // fitModel.R:181-183
    for (int k = 0; k < max_cycles; k++) em_ll[k] = NAN;
    PhaseTimer pt("fit_model");
    {   // initial values; the warm-start matrices survive the initialisation
        std::vector<double> keep[4];
        double *mats[4] = {p->mat_mu, p->batch_mu, p->mat_disp, p->batch_disp};
        const size_t sizes[4] = {(size_t)G * M.p_mu, (size_t)G * M.nbf_mu, (size_t)G * M.p_disp, (size_t)G * M.nbf_disp};
        for (int k = 0; k < 4; k++)
            if ((keep_mask >> k & 1) && mats[k]) keep[k].assign(mats[k], mats[k] + sizes[k]);
        if ((rc = init_params_impl(ctx, M, p, min_rowmean_global, fit_drop != 0))) return rc;
        for (int k = 0; k < 4; k++)
            if (!keep[k].empty()) memcpy(mats[k], keep[k].data(), sizeof(double) * sizes[k]);
    }
    pt.lap("initial values");
    if ((rc = eval_loglik_impl(ctx, M, p, ll))) return rc;  // :228-233
    double ll_new, ll_old = NAN;
    if ((rc = sum_ll(ctx, ll, G, &ll_new))) return rc;
    if (ll_init) *ll_init = ll_new;
    pt.lap("initial log-likelihood");
    int iter = 1;
    std::vector<double> pill(fit_drop ? N : 1);
    std::vector<int32_t> piconv(fit_drop ? N : 1);
    int any_nonconv = 0;
    while (iter == 1 || (ll_new > ll_old * c->prec_em && iter <= max_cycles)) {  // :242-243
        if (fit_drop && (rc = fit_pi_impl(ctx, M, m->drop_fit_group, p, c, pill.data(), piconv.data()))) return rc;
        pt.lap("fitPi");
        if ((rc = sync_error(ctx, fit_mudisp_impl(ctx, M, p, c, ll, conv)))) return rc;
        pt.lap("fitMuDisp");
        ll_old = ll_new;
        if ((rc = sum_ll(ctx, ll, G, &ll_new))) return rc;
        em_ll[iter - 1] = ll_new;
        iter++;
    }
    for (int i = 0; i < G; i++)
        if (conv[i] != 0) any_nonconv = 1;
    if (ctx->world > 1) {  // all(vecMuDispEstConverged == 0) over every shard
        double v = any_nonconv;
        CK(cudaMemcpyAsync(ctx->d_xchg, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = do_allreduce(ctx, ctx->d_xchg, 1, LPB200_F64))) return rc;
        CK(cudaMemcpyAsync(&v, ctx->d_xchg, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        any_nonconv = v > 0;
    }
    *converged = (!any_nonconv && ll_new < ll_old * c->prec_em && ll_new > ll_old) ? 1 : 0;  // :342-347
    *n_iter = iter - 1;
    return 0;
}

// Several independent fitModel() calls without drop-out estimation (the fits B, C, D of fitLPModels,
// fitWrapperLP.R:152-238: each is initialisation + one fitMuDisp pass) run concurrently: their fit
// kernels go to separate streams, so the persistent CTAs of one kernel take over the SMs that the
// item-length tail of another leaves idle.  Results are identical to n_jobs sequential
// lpb200_fit_model(fit_drop = 0) calls.
extern "C" int lpb200_fit_model_multi(lpb200_ctx *ctx, int n_jobs, const lpb200_model *models, lpb200_params *params,
                                      const lpb200_control *c, double min_rowmean_global, double *ll_init, double *em_ll,
                                      int32_t *converged, double *ll, int32_t *conv, uint64_t *ole) {
    if (is_multi(ctx)) return multi_fit_model_multi(ctx, n_jobs, models, params, c, min_rowmean_global, ll_init, em_ll, converged, ll, conv, ole);
    int rc = require_ready(ctx);
    if (rc) return rc;
    enum { MAX_JOBS = 4 };
    if (n_jobs < 1 || n_jobs > MAX_JOBS) return set_err(ctx, LPB200_E_ARG, "fit_model_multi: 1..%d jobs", (int)MAX_JOBS);
    if (!models || !params || !ll_init || !em_ll || !converged || !ll || !conv)
        return set_err(ctx, LPB200_E_ARG, "fit_model_multi: NULL argument");
    lpb200_control ctl_default;
    c = control_or_default(c, ctl_default);
    const int G = ctx->G;
    FitJob jobs[MAX_JOBS];
    for (int j = 0; j < n_jobs; j++) {
        FitJob &J = jobs[j];
        if ((rc = make_model(ctx, &models[j], J.M))) return rc;
        J.p = &params[j];
        J.ll = ll + (size_t)j * G;
        J.conv = conv + (size_t)j * G;
        if ((rc = check_params(ctx, J.M, J.p, true))) return rc;
        if ((rc = init_params_impl(ctx, J.M, J.p, min_rowmean_global, false))) return rc;
        if ((rc = eval_loglik_impl(ctx, J.M, J.p, J.ll))) return rc;  // fitModel.R:228-233
        if ((rc = sum_ll(ctx, J.ll, G, &ll_init[j]))) return rc;
    }
    while ((int)ctx->aux_streams.size() < n_jobs - 1) {
        cudaStream_t st;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ctx->aux_streams.push_back(st);
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));  // everything enqueued so far is visible to the auxiliary streams
    // Launch the expensive fits first (impulse: 3 starts x 8 parameters) and the cheap ones last: a later
    // kernel's CTAs become resident as the earlier kernel's persistent CTAs retire, so the cheap fit is
    // what fills the item-length tail of the last expensive one.  The order does not change any result.
    int order[MAX_JOBS];
    for (int j = 0; j < n_jobs; j++) order[j] = j;
    auto cost = [&](int j) { return (jobs[j].M.mu_model == LPB200_MU_IMPULSE ? 3 : 1) * jobs[j].M.n_theta; };
    std::stable_sort(order, order + n_jobs, [&](int a, int b) { return cost(a) > cost(b); });
    for (int q = 0; q < n_jobs; q++) {
        const int j = order[q];
        jobs[j].stream = (q == 0) ? ctx->stream : ctx->aux_streams[q - 1];
        if (q > 0) CK(cudaStreamWaitEvent(jobs[j].stream, ctx->ev0, 0));
        if ((rc = fit_launch(ctx, jobs[j], c))) {
            cudaDeviceSynchronize();
            return rc;
        }
    }
    int first_err = 0;
    float span = 0.f;
    for (int j = 0; j < n_jobs; j++) {
        rc = fit_collect(ctx, jobs[j], nullptr);
        if (rc && !first_err) first_err = rc;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, jobs[order[0]].e0, jobs[j].e1);
        span = std::max(span, ms);
        if (getenv("LPB200_DEBUG")) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, jobs[order[0]].e0, jobs[j].e0);
            fprintf(stderr, "[lpb200] concurrent job %d: kernel from %.1f ms to %.1f ms\n", j, t0, ms);
        }
        if (ole) ole[j] = jobs[j].ole;
    }
    ctx->stats.kernel_ms += span;
    if ((first_err = sync_error(ctx, first_err))) return first_err;
    for (int j = 0; j < n_jobs; j++) {
        double ll_new;
        if ((rc = sum_ll(ctx, jobs[j].ll, G, &ll_new))) return rc;
        em_ll[j] = ll_new;
        int any_nonconv = 0;
        for (int i = 0; i < G; i++)
            if (jobs[j].conv[i] != 0) any_nonconv = 1;
        if (ctx->world > 1) {
            double v = any_nonconv;
            CK(cudaMemcpyAsync(ctx->d_xchg, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            if ((rc = do_allreduce(ctx, ctx->d_xchg, 1, LPB200_F64))) return rc;
            CK(cudaMemcpyAsync(&v, ctx->d_xchg, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            any_nonconv = v > 0;
        }
        converged[j] = (!any_nonconv && ll_new < ll_init[j] * c->prec_em && ll_new > ll_init[j]) ? 1 : 0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// calcPostDrop_Matrix
// ---------------------------------------------------------------------------------------------
extern "C" int lpb200_expand_fits(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, const int32_t *gene_ids,
                                  int n_ids, int what, double *z) {
    if (is_multi(ctx)) return multi_expand_fits(ctx, m, p, gene_ids, n_ids, what, z);
    int rc = require_ready(ctx);
    if (rc) return rc;
    DevModel M;
    if ((rc = make_model(ctx, m, M))) return rc;
    const int kind = what & 15;
    if (kind > LPB200_EXPAND_NORMDATA) return set_err(ctx, LPB200_E_ARG, "unknown expansion %d", what);
    const bool needs_drop = kind == LPB200_EXPAND_DROP || kind == LPB200_EXPAND_POSTDROP;
    if (needs_drop && (M.drop_model == LPB200_DROP_NONE || !p->mat_drop))
        return set_err(ctx, LPB200_E_ARG, "this expansion needs a drop-out model");
    if ((rc = check_params(ctx, M, p, false))) return rc;
    if (n_ids <= 0) return 0;
    if (!gene_ids || !z) return set_err(ctx, LPB200_E_ARG, "expand_fits: NULL argument");
    const int G = ctx->G, N = ctx->N;
    for (int r = 0; r < n_ids; r++)
        if (gene_ids[r] < 0 || gene_ids[r] >= G) return set_err(ctx, LPB200_E_ARG, "gene id out of range");
    std::vector<double> nat;
    params_to_nat(ctx, M, p, nat);
    DevBuf<double> d_nat, d_drop, d_z;
    DevBuf<int32_t> d_ids;
    CK(d_nat.alloc(nat.size()));
    CK(d_drop.alloc((size_t)N * std::max(1, M.q_drop)));
    CK(d_z.alloc((size_t)n_ids * N));
    CK(d_ids.alloc(n_ids));
    CK(cudaMemcpyAsync(d_nat.p, nat.data(), sizeof(double) * nat.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (needs_drop)
        CK(cudaMemcpyAsync(d_drop.p, p->mat_drop, sizeof(double) * (size_t)N * M.q_drop, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_ids.p, gene_ids, sizeof(int32_t) * n_ids, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((N + 255) / 256, n_ids);
    if (ctx->is_u16) expand_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(ctx->D, M, count_view(ctx), d_nat.p, d_drop.p, d_ids.p, n_ids, what, d_z.p);
    else expand_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>(ctx->D, M, count_view(ctx), d_nat.p, d_drop.p, d_ids.p, n_ids, what, d_z.p);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    CK(cudaMemcpyAsync(z, d_z.p, sizeof(double) * (size_t)n_ids * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lpb200_post_drop(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, const int32_t *gene_ids,
                                int n_ids, double *z) {
    return lpb200_expand_fits(ctx, m, p, gene_ids, n_ids, LPB200_EXPAND_POSTDROP, z);
}

// ---------------------------------------------------------------------------------------------
// LRT (runHypothesisTests.R:81-86): pchisq upper tail + Benjamini-Hochberg.  Host arithmetic on
// G doubles (negligible); regularised upper incomplete gamma by series / Lentz
// continued fraction in long double.
// ---------------------------------------------------------------------------------------------
static double gamma_q(long double a, long double x) {
    if (x <= 0) return 1.0;
    const long double lg = lgammal(a);
    if (x < a + 1.0L) {  // P by series, Q = 1 - P
        long double ap = a, sum = 1.0L / a, del = sum;
        for (int n = 0; n < 100000; n++) {
            ap += 1.0L;
            del *= x / ap;
            sum += del;
            if (fabsl(del) < fabsl(sum) * 1e-19L) break;
        }
        long double P = sum * expl(-x + a * logl(x) - lg);
        return (double)(1.0L - P);
    }
    const long double tiny = 1e-4000L;
    long double b = x + 1.0L - a, cc = 1.0L / tiny, d = 1.0L / b, h = d;
    for (int i = 1; i < 100000; i++) {
        long double an = -(long double)i * ((long double)i - a);
        b += 2.0L;
        d = an * d + b;
        if (fabsl(d) < tiny) d = tiny;
        cc = b + an / cc;
        if (fabsl(cc) < tiny) cc = tiny;
        d = 1.0L / d;
        long double del = d * cc;
        h *= del;
        if (fabsl(del - 1.0L) < 1e-19L) break;
    }
    return (double)(expl(-x + a * logl(x) - lg) * h);
}

static double pchisq_upper(double q, double df) {
    if (std::isnan(q) || std::isnan(df) || df < 0) return NAN;
    if (q <= 0) return 1.0;
    if (df == 0) return 0.0;
    if (std::isinf(q)) return 0.0;
    return gamma_q(0.5L * df, 0.5L * q);
}

extern "C" int lpb200_lrt(const double *ll_full, const double *ll_red, int ddf, int G, double *p, double *padj) {
    if (!ll_full || !ll_red || !p || !padj || G < 0) return LPB200_E_ARG;
    std::vector<int> idx;
    idx.reserve(G);
    for (int i = 0; i < G; i++) {
        double dev = 2 * (ll_full[i] - ll_red[i]);
        p[i] = pchisq_upper(dev, (double)ddf);
        padj[i] = NAN;
        if (!std::isnan(p[i])) idx.push_back(i);
    }
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return p[a] > p[b]; });
    const int n = (int)idx.size();
    double cm = INFINITY;
    for (int k = 0; k < n; k++) {  // p.adjust BH: pmin(1, cummin(n / i * p[o])), i = n..1
        double v = (double)n / (double)(n - k) * p[idx[k]];
        if (v < cm) cm = v;
        padj[idx[k]] = cm < 1.0 ? cm : 1.0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// FP64 DFMA peak probe
// ---------------------------------------------------------------------------------------------
extern "C" int lpb200_fp64_peak(lpb200_ctx *ctx, double *tflops) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) ctx = ctx->shards[0];
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> out;
    CK(out.alloc((size_t)blocks * threads));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out.p, 1024);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out.p, iters);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CK(cudaGetLastError());
    ctx->stats.launches += 4;
    *tflops = best;
    return 0;
}

extern "C" int lpb200_libm_probe(lpb200_ctx *ctx, const double *x, int n, int use_library, double *exp_out, double *log_out) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) ctx = ctx->shards[0];
    if (!x || !exp_out || !log_out || n <= 0) return set_err(ctx, LPB200_E_ARG, "libm_probe: bad arguments");
    CK(cudaSetDevice(ctx->device));
    DevBuf<double> dx, de, dl;
    CK(dx.alloc(n));
    CK(de.alloc(n));
    CK(dl.alloc(n));
    CK(cudaMemcpyAsync(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    libm_probe_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(dx.p, n, use_library, de.p, dl.p);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    CK(cudaMemcpyAsync(exp_out, de.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(log_out, dl.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lpb200_div_probe(lpb200_ctx *ctx, const double *a, const double *b, int n, int use_library, double *quot_out,
                                double *rcp_out) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) ctx = ctx->shards[0];
    if (!a || !b || !quot_out || !rcp_out || n <= 0) return set_err(ctx, LPB200_E_ARG, "div_probe: bad arguments");
    CK(cudaSetDevice(ctx->device));
    DevBuf<double> da, db, dq, dr;
    CK(da.alloc(n));
    CK(db.alloc(n));
    CK(dq.alloc(n));
    CK(dr.alloc(n));
    CK(cudaMemcpyAsync(da.p, a, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(db.p, b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    div_probe_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(da.p, db.p, n, use_library, dq.p, dr.p);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    CK(cudaMemcpyAsync(quot_out, dq.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(rcp_out, dr.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int lpb200_rcp64h_probe(lpb200_ctx *ctx, const int32_t *hi, int n, int32_t lo, int32_t *out_hi, int32_t *out_lo) {
    if (!ctx) return LPB200_E_CUDA;
    if (is_multi(ctx)) ctx = ctx->shards[0];
    if (!hi || !out_hi || !out_lo || n <= 0) return set_err(ctx, LPB200_E_ARG, "rcp64h_probe: bad arguments");
    CK(cudaSetDevice(ctx->device));
    DevBuf<int32_t> dh, doh, dol;
    CK(dh.alloc(n));
    CK(doh.alloc(n));
    CK(dol.alloc(n));
    CK(cudaMemcpyAsync(dh.p, hi, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    rcp64h_probe_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(dh.p, n, lo, doh.p, dol.p);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    CK(cudaMemcpyAsync(out_hi, doh.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_lo, dol.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

#include "lp_multi.cuh"
