// lp_multi.cuh -- single-process multi-device contexts (included at the end of lp_api.cu).
//
// lpb200_create_multi(devices, n) returns an lpb200_ctx that owns one shard context per device, an NCCL
// communicator over them (ncclCommInitAll) and one stream per device.  Every entry point of include/lpb200.h
// accepts such a context: the genes are dealt to the shards in contiguous ranges starting on multiples of the
// summation block (LP_PI_BLOCK = 16 genes, so results are bit-identical to a single-device run), the per-gene
// arrays of the call are sliced / gathered on the host, and one host thread per device drives its shard through
// the ordinary single-device code path.  The shards meet in the NCCL all-reduces of fitPi's lock-step rounds and
// of the EM log-likelihood.  This is the replacement of the reference's BiocParallel pool
// (main_LineagePulse.R:271-275: MulticoreParam(workers = scaNProc)) that the public API reaches.
#pragma once
#include <chrono>
#include <thread>

static bool is_multi(const lpb200_ctx *c) { return c && !c->shards.empty(); }

// NCCL communicators are expensive to create (ncclCommInitAll: 0.5-3 s for 2-8 GPUs, plus the lazy connection set-up
// of the first collective), contexts are cheap and the public API makes one per count matrix: communicators are
// therefore kept per device list for the life of the process and lent to one context at a time.
struct CommSet {
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    bool in_use = false;
};
static std::vector<CommSet *> g_comm_sets;
static std::mutex g_comm_mutex;

static CommSet *acquire_comms(const int *devices, int n) {
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    std::vector<int> key(devices, devices + n);
    for (CommSet *cs : g_comm_sets)
        if (!cs->in_use && cs->devices == key) {
            cs->in_use = true;
            return cs;
        }
    NcclApi *nccl = nccl_api();
    if (!nccl) return nullptr;
    CommSet *cs = new CommSet();
    cs->devices = key;
    cs->comms.assign(n, nullptr);
    if (nccl->CommInitAll(cs->comms.data(), n, devices) != ncclSuccess) {
        delete cs;
        return nullptr;
    }
    // connection set-up happens at a communicator's first collective: pay it here, once per device list
    {
        std::vector<double *> bufs(n, nullptr);
        bool ok = true;
        for (int k = 0; k < n && ok; k++)
            ok = cudaSetDevice(devices[k]) == cudaSuccess && cudaMalloc((void **)&bufs[k], sizeof(double)) == cudaSuccess &&
                 cudaMemset(bufs[k], 0, sizeof(double)) == cudaSuccess;
        if (ok) {
            NcclGroup group(nccl);
            for (int k = 0; k < n; k++) {
                cudaSetDevice(devices[k]);
                nccl->AllReduce(bufs[k], bufs[k], 1, ncclFloat64, ncclSum, cs->comms[k], (cudaStream_t)0);
            }
        }
        for (int k = 0; k < n; k++) {
            cudaSetDevice(devices[k]);
            cudaDeviceSynchronize();
            if (bufs[k]) cudaFree(bufs[k]);
        }
    }
    cs->in_use = true;
    g_comm_sets.push_back(cs);
    return cs;
}
static void release_comms(CommSet *cs) {
    if (!cs) return;
    std::lock_guard<std::mutex> lock(g_comm_mutex);
    cs->in_use = false;
}

// contiguous gene ranges of near-equal size whose starts are multiples of LP_PI_BLOCK (as distributed.shard_bounds)
static std::vector<int> shard_bounds(int G, int n) {
    const int n_blocks = (G + LP_PI_BLOCK - 1) / LP_PI_BLOCK;
    const int base = n_blocks / n, extra = n_blocks % n;
    std::vector<int> b(n + 1, 0);
    for (int r = 0; r < n; r++) b[r + 1] = std::min(G, b[r] + (base + (r < extra ? 1 : 0)) * LP_PI_BLOCK);
    b[n] = G;
    return b;
}

extern "C" lpb200_ctx *lpb200_create_multi(const int *devices, int n) {
    if (!devices || n < 1 || n > 64) return nullptr;
    if (n == 1) return lpb200_create(devices[0]);
    bool repeated = false, all_same = true;
    for (int a = 0; a < n; a++) {
        if (devices[a] != devices[0]) all_same = false;
        for (int b = a + 1; b < n; b++)
            if (devices[a] == devices[b]) repeated = true;
    }
    // distinct devices: NCCL.  One device n times (n <= 8): the loopback communicator (lp_comm.cuh; testing).
    if (repeated && (!all_same || n > 8)) return nullptr;
    const bool debug = getenv("LPB200_DEBUG") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!debug) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[lpb200] create_multi: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    CommSet *cs = nullptr;
    if (!repeated) {
        cs = acquire_comms(devices, n);
        if (!cs) return nullptr;
    }
    lap("communicator");
    lpb200_ctx *parent = new lpb200_ctx();
    parent->device = devices[0];
    parent->comm_set = cs;
    bool ok = true;
    for (int k = 0; k < n && ok; k++) {
        lpb200_ctx *c = lpb200_create(devices[k]);
        if (!c) { ok = false; break; }
        parent->shards.push_back(c);
        if (cudaSetDevice(devices[k]) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) ok = false;
        else c->owns_stream = true;
    }
    lap("shard contexts");
    if (!ok) {
        release_comms(cs);
        for (lpb200_ctx *c : parent->shards) lpb200_destroy(c);
        delete parent;
        return nullptr;
    }
    Loopback *L = nullptr;
    if (repeated) {
        L = new Loopback();
        L->n = n;
        L->bufs.assign(n, nullptr);
        L->tmp.assign(n, nullptr);
        L->tmp_words.assign(n, 0);
    }
    for (int k = 0; k < n; k++) {
        parent->shards[k]->comm = cs ? cs->comms[k] : nullptr;
        parent->shards[k]->comm_borrowed = true;
        parent->shards[k]->loopback = L;
        parent->shards[k]->rank = k;
        parent->shards[k]->world = n;
    }
    parent->sm_count = parent->shards[0]->sm_count;
    return parent;
}

extern "C" int lpb200_n_shards(lpb200_ctx *ctx) { return !ctx ? 0 : (is_multi(ctx) ? (int)ctx->shards.size() : 1); }

extern "C" int lpb200_shard_bounds(lpb200_ctx *ctx, int32_t *bounds) {
    if (!ctx || !bounds) return LPB200_E_ARG;
    if (!is_multi(ctx)) {
        bounds[0] = 0;
        bounds[1] = ctx->G;
        return 0;
    }
    if (ctx->bounds.empty()) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
    for (size_t k = 0; k < ctx->bounds.size(); k++) bounds[k] = ctx->bounds[k];
    return 0;
}

// one host thread per shard; the first failing shard's code and message become the parent's
template <typename F>
static int multi_run(lpb200_ctx *ctx, F &&fn) {
    const int n = (int)ctx->shards.size();
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int k = 1; k < n; k++) th.emplace_back([&, k]() { rcs[k] = fn(k, ctx->shards[k]); });
    rcs[0] = fn(0, ctx->shards[0]);
    for (auto &t : th) t.join();
    for (int k = 0; k < n; k++)
        if (rcs[k] != 0) return set_err(ctx, rcs[k], "shard %d (device %d): %s", k, ctx->shards[k]->device, ctx->shards[k]->err.c_str());
    return 0;
}

static int multi_require_counts(lpb200_ctx *ctx) {
    if (ctx->bounds.empty()) return set_err(ctx, LPB200_E_STATE, "counts not loaded");
    return 0;
}

// ---- slicing of R matrices (column-major, leading dimension G) by gene range --------------------------------
static void slice_cols(const double *full, int G, int g0, int gs, int ncols, std::vector<double> &out) {
    out.resize((size_t)gs * std::max(ncols, 0));
    for (int k = 0; k < ncols; k++) memcpy(out.data() + (size_t)k * gs, full + (size_t)k * G + g0, sizeof(double) * gs);
}
static void unslice_cols(double *full, int G, int g0, int gs, int ncols, const std::vector<double> &in) {
    for (int k = 0; k < ncols; k++) memcpy(full + (size_t)k * G + g0, in.data() + (size_t)k * gs, sizeof(double) * gs);
}

struct ShardParams {
    std::vector<double> mu, bmu, disp, bdisp, drop;
    lpb200_params p{};
};

// sizes of the parameter matrices of model m under the loaded design
static int multi_model(lpb200_ctx *ctx, const lpb200_model *m, DevModel &M) {
    int rc = multi_require_counts(ctx);
    if (rc) return rc;
    if (!m) return set_err(ctx, LPB200_E_ARG, "model is NULL");
    lpb200_ctx *c0 = ctx->shards[0];
    if (!c0->have_design) return set_err(ctx, LPB200_E_STATE, "design not set");
    if ((rc = make_model(c0, m, M))) return set_err(ctx, rc, "%s", c0->err.c_str());
    return 0;
}

static void slice_params(lpb200_ctx *ctx, const DevModel &M, const lpb200_params *full, int k, ShardParams &sp, bool with_drop) {
    const int G = ctx->G, N = ctx->N, g0 = ctx->bounds[k], gs = ctx->bounds[k + 1] - g0;
    sp.p = lpb200_params{};
    if (full->mat_mu) { slice_cols(full->mat_mu, G, g0, gs, M.p_mu, sp.mu); sp.p.mat_mu = sp.mu.data(); }
    if (full->mat_disp) { slice_cols(full->mat_disp, G, g0, gs, M.p_disp, sp.disp); sp.p.mat_disp = sp.disp.data(); }
    if (full->batch_mu && M.nbf_mu > 0) { slice_cols(full->batch_mu, G, g0, gs, M.nbf_mu, sp.bmu); sp.p.batch_mu = sp.bmu.data(); }
    if (full->batch_disp && M.nbf_disp > 0) { slice_cols(full->batch_disp, G, g0, gs, M.nbf_disp, sp.bdisp); sp.p.batch_disp = sp.bdisp.data(); }
    if (with_drop && full->mat_drop && M.q_drop > 0) {   // per-cell: replicated on every shard
        sp.drop.assign(full->mat_drop, full->mat_drop + (size_t)N * M.q_drop);
        sp.p.mat_drop = sp.drop.data();
    }
}
static void unslice_params(lpb200_ctx *ctx, const DevModel &M, lpb200_params *full, int k, const ShardParams &sp, bool with_drop) {
    const int G = ctx->G, N = ctx->N, g0 = ctx->bounds[k], gs = ctx->bounds[k + 1] - g0;
    if (full->mat_mu) unslice_cols(full->mat_mu, G, g0, gs, M.p_mu, sp.mu);
    if (full->mat_disp) unslice_cols(full->mat_disp, G, g0, gs, M.p_disp, sp.disp);
    if (full->batch_mu && M.nbf_mu > 0) unslice_cols(full->batch_mu, G, g0, gs, M.nbf_mu, sp.bmu);
    if (full->batch_disp && M.nbf_disp > 0) unslice_cols(full->batch_disp, G, g0, gs, M.nbf_disp, sp.bdisp);
    // the drop-out model is identical on every shard (exact cross-shard sums): shard 0's copy is the result
    if (with_drop && k == 0 && full->mat_drop && M.q_drop > 0) memcpy(full->mat_drop, sp.drop.data(), sizeof(double) * (size_t)N * M.q_drop);
}

// ---- data -----------------------------------------------------------------------------------------------------
static int multi_begin_load(lpb200_ctx *ctx, int G, int N) {
    const int n = (int)ctx->shards.size();
    if (G < n) return set_err(ctx, LPB200_E_ARG, "fewer genes (%d) than devices (%d)", G, n);
    ctx->bounds = shard_bounds(G, n);
    for (int k = 0; k < n; k++)
        if (ctx->bounds[k + 1] <= ctx->bounds[k]) return set_err(ctx, LPB200_E_ARG, "%d genes cannot be dealt to %d devices in blocks of %d", G, n, LP_PI_BLOCK);
    ctx->G = G;
    ctx->N = N;
    return 0;
}

static int load_counts_dense_ld(lpb200_ctx *ctx, int G, int N, const double *x, size_t ld);

static int multi_load_counts_dense(lpb200_ctx *ctx, int G, int N, const double *x) {
    if (G <= 0 || N <= 0 || !x) return set_err(ctx, LPB200_E_ARG, "load_counts_dense: bad arguments");
    int rc = multi_begin_load(ctx, G, N);
    if (rc) return rc;
    rc = multi_run(ctx, [&](int k, lpb200_ctx *c) { return load_counts_dense_ld(c, ctx->bounds[k + 1] - ctx->bounds[k], N, x + ctx->bounds[k], (size_t)G); });
    if (rc) ctx->bounds.clear();
    return rc;
}

static int multi_load_counts_i32(lpb200_ctx *ctx, int G, int N, const int32_t *x, int is_device) {
    if (G <= 0 || N <= 0 || !x) return set_err(ctx, LPB200_E_ARG, "load_counts_i32: bad arguments");
    if (is_device) return set_err(ctx, LPB200_E_ARG, "load_counts_i32: a multi-device context takes host counts (each shard uploads its own rows)");
    int rc = multi_begin_load(ctx, G, N);
    if (rc) return rc;
    rc = multi_run(ctx, [&](int k, lpb200_ctx *c) {
        return lpb200_load_counts_i32(c, ctx->bounds[k + 1] - ctx->bounds[k], N, x + (size_t)ctx->bounds[k] * N, 0);
    });
    if (rc) ctx->bounds.clear();
    return rc;
}

static int multi_load_counts_csc(lpb200_ctx *ctx, int G, int N, const int32_t *p, const int32_t *i, const double *x) {
    if (G <= 0 || N <= 0 || !p) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: bad arguments");
    if (p[0] != 0 || p[N] < 0 || (p[N] > 0 && (!i || !x))) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: bad column pointers");
    for (int j = 0; j < N; j++)
        if (p[j + 1] < p[j]) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: column pointers decrease at column %d", j);
    for (int32_t e = 0; e < p[N]; e++)
        if (i[e] < 0 || i[e] >= G) return set_err(ctx, LPB200_E_ARG, "load_counts_csc: row index %d out of range", (int)i[e]);
    int rc = multi_begin_load(ctx, G, N);
    if (rc) return rc;
    rc = multi_run(ctx, [&](int k, lpb200_ctx *c) {   // the shard's rows of the dgCMatrix
        const int g0 = ctx->bounds[k], g1 = ctx->bounds[k + 1];
        std::vector<int32_t> sp(N + 1, 0), si;
        std::vector<double> sx;
        for (int j = 0; j < N; j++) {
            for (int32_t e = p[j]; e < p[j + 1]; e++)
                if (i[e] >= g0 && i[e] < g1) {
                    si.push_back(i[e] - g0);
                    sx.push_back(x[e]);
                }
            sp[j + 1] = (int32_t)si.size();
        }
        return lpb200_load_counts_csc(c, g1 - g0, N, sp.data(), si.data(), sx.data());
    });
    if (rc) ctx->bounds.clear();
    return rc;
}

static int multi_set_design(lpb200_ctx *ctx, const lpb200_design *d) {
    int rc = multi_require_counts(ctx);
    if (rc) return rc;
    if (!d) return set_err(ctx, LPB200_E_ARG, "design is NULL");
    if (d->n_genes != ctx->G || d->n_cells != ctx->N)
        return set_err(ctx, LPB200_E_ARG, "design is %d x %d but counts are %d x %d", d->n_genes, d->n_cells, ctx->G, ctx->N);
    return multi_run(ctx, [&](int k, lpb200_ctx *c) {
        const int g0 = ctx->bounds[k], gs = ctx->bounds[k + 1] - g0;
        lpb200_design ds = *d;
        ds.n_genes = gs;
        std::vector<double> pred;
        if (d->pi_const_pred && d->n_pi_pred > 0) {
            slice_cols(d->pi_const_pred, ctx->G, g0, gs, d->n_pi_pred, pred);
            ds.pi_const_pred = pred.data();
        }
        return lpb200_set_design(c, &ds);
    });
}

static double multi_min_rowmean(lpb200_ctx *ctx, double given) {
    if (!std::isnan(given)) return given;
    double m = INFINITY;
    for (lpb200_ctx *c : ctx->shards)
        for (double v : c->row_means) m = std::min(m, v);
    return m;
}

static int multi_row_means(lpb200_ctx *ctx, double *means) {
    int rc = multi_require_counts(ctx);
    if (rc) return rc;
    for (size_t k = 0; k < ctx->shards.size(); k++)
        memcpy(means + ctx->bounds[k], ctx->shards[k]->row_means.data(), sizeof(double) * ctx->shards[k]->G);
    return 0;
}

// ---- compute --------------------------------------------------------------------------------------------------
static int multi_init_params(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *out, double min_rowmean_global) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (!out) return set_err(ctx, LPB200_E_ARG, "params is NULL");
    const double mrm = multi_min_rowmean(ctx, min_rowmean_global);
    std::vector<ShardParams> sp(ctx->shards.size());
    rc = multi_run(ctx, [&](int k, lpb200_ctx *c) {
        slice_params(ctx, M, out, k, sp[k], true);
        return lpb200_init_params(c, m, &sp[k].p, mrm);
    });
    if (rc) return rc;
    for (size_t k = 0; k < sp.size(); k++) unslice_params(ctx, M, out, (int)k, sp[k], true);
    return 0;
}

static int multi_eval_loglik(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, double *ll) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (!p || !ll) return set_err(ctx, LPB200_E_ARG, "eval_loglik: NULL argument");
    return multi_run(ctx, [&](int k, lpb200_ctx *c) {
        ShardParams sp;
        slice_params(ctx, M, p, k, sp, true);
        return lpb200_eval_loglik(c, m, &sp.p, ll + ctx->bounds[k]);
    });
}

static int multi_impulse_starts(lpb200_ctx *ctx, double *peak, double *valley) {
    int rc = multi_require_counts(ctx);
    if (rc) return rc;
    const int G = ctx->G;
    return multi_run(ctx, [&](int k, lpb200_ctx *c) {
        const int g0 = ctx->bounds[k], gs = ctx->bounds[k + 1] - g0;
        std::vector<double> pk((size_t)gs * 7), vl((size_t)gs * 7);
        int r = lpb200_impulse_starts(c, pk.data(), vl.data());
        if (r) return r;
        unslice_cols(peak, G, g0, gs, 7, pk);
        unslice_cols(valley, G, g0, gs, 7, vl);
        return 0;
    });
}

static int multi_fit_mudisp_items(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll,
                                  int32_t *conv, int32_t *n_starts, int32_t *n_theta, double *theta0, double *theta,
                                  double *ll_items, int32_t *conv_items, uint32_t *evals_items) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (!p || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_mudisp: NULL argument");
    const int ns = (M.mu_model == LPB200_MU_IMPULSE) ? 3 : 1, n = M.n_theta;
    if (n_starts) *n_starts = ns;
    if (n_theta) *n_theta = n;
    std::vector<ShardParams> sp(ctx->shards.size());
    rc = multi_run(ctx, [&](int k, lpb200_ctx *sc) {
        slice_params(ctx, M, p, k, sp[k], true);
        const size_t i0 = (size_t)ctx->bounds[k] * ns;   // items are gene-major: the shard's items are contiguous
        int32_t ns_k, nt_k;
        return lpb200_fit_mudisp_items(sc, m, &sp[k].p, c, ll + ctx->bounds[k], conv + ctx->bounds[k], &ns_k, &nt_k,
                                       theta0 ? theta0 + i0 * n : nullptr, theta ? theta + i0 * n : nullptr,
                                       ll_items ? ll_items + i0 : nullptr, conv_items ? conv_items + i0 : nullptr,
                                       evals_items ? evals_items + i0 : nullptr);
    });
    // parameters of the shards that succeeded are written back either way (as a single-device call does)
    for (size_t k = 0; k < sp.size(); k++)
        if (!sp[k].mu.empty()) unslice_params(ctx, M, p, (int)k, sp[k], false);
    return rc;
}

static int multi_fit_pi(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, const lpb200_control *c, double *ll, int32_t *conv) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (!p || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_pi: NULL argument");
    const size_t n_out = (m->drop_fit_group == LPB200_DROPFIT_PERCELL) ? (size_t)ctx->N : 1;
    std::vector<ShardParams> sp(ctx->shards.size());
    rc = multi_run(ctx, [&](int k, lpb200_ctx *sc) {
        slice_params(ctx, M, p, k, sp[k], true);
        if (k == 0) return lpb200_fit_pi(sc, m, &sp[k].p, c, ll, conv);
        std::vector<double> ll_k(n_out);          // identical on every shard; shard 0's copy is returned
        std::vector<int32_t> conv_k(n_out);
        return lpb200_fit_pi(sc, m, &sp[k].p, c, ll_k.data(), conv_k.data());
    });
    if (!sp[0].drop.empty() && p->mat_drop) memcpy(p->mat_drop, sp[0].drop.data(), sizeof(double) * sp[0].drop.size());
    return rc;
}

static int multi_fit_model(lpb200_ctx *ctx, const lpb200_model *m, lpb200_params *p, int fit_drop, const lpb200_control *c,
                           double min_rowmean_global, int keep_mask, double *em_ll, int32_t *n_iter, int32_t *converged,
                           double *ll_init, double *ll, int32_t *conv) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (!p || !em_ll || !n_iter || !converged || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_model: NULL argument");
    lpb200_control ctl_default;
    c = control_or_default(c, ctl_default);
    const double mrm = multi_min_rowmean(ctx, min_rowmean_global);
    const int n = (int)ctx->shards.size(), cyc = std::max(1, c->max_cycles);
    std::vector<ShardParams> sp(n);
    std::vector<std::vector<double>> em(n, std::vector<double>(cyc, NAN));
    std::vector<int32_t> it(n, 0), cv(n, 0);
    std::vector<double> li(n, NAN);
    rc = multi_run(ctx, [&](int k, lpb200_ctx *sc) {
        slice_params(ctx, M, p, k, sp[k], true);
        return lpb200_fit_model_from(sc, m, &sp[k].p, fit_drop, c, mrm, keep_mask, em[k].data(), &it[k], &cv[k], &li[k],
                                     ll + ctx->bounds[k], conv + ctx->bounds[k]);
    });
    if (rc) return rc;
    for (int k = 0; k < n; k++) unslice_params(ctx, M, p, k, sp[k], true);
    const int max_cycles = (fit_drop && M.drop_model != LPB200_DROP_NONE) ? c->max_cycles : 1;
    for (int q = 0; q < max_cycles; q++) em_ll[q] = em[0][q];   // the EM trajectory is the same on every shard
    *n_iter = it[0];
    *converged = cv[0];
    if (ll_init) *ll_init = li[0];
    return 0;
}

static int multi_fit_model_multi(lpb200_ctx *ctx, int n_jobs, const lpb200_model *models, lpb200_params *params,
                                 const lpb200_control *c, double min_rowmean_global, double *ll_init, double *em_ll,
                                 int32_t *converged, double *ll, int32_t *conv, uint64_t *ole) {
    int rc = multi_require_counts(ctx);
    if (rc) return rc;
    if (n_jobs < 1 || n_jobs > 4) return set_err(ctx, LPB200_E_ARG, "fit_model_multi: 1..4 jobs");
    if (!models || !params || !ll_init || !em_ll || !converged || !ll || !conv) return set_err(ctx, LPB200_E_ARG, "fit_model_multi: NULL argument");
    DevModel M[4];
    for (int j = 0; j < n_jobs; j++)
        if ((rc = multi_model(ctx, &models[j], M[j]))) return rc;
    const double mrm = multi_min_rowmean(ctx, min_rowmean_global);
    const int n = (int)ctx->shards.size(), G = ctx->G;
    std::vector<std::vector<ShardParams>> sp(n, std::vector<ShardParams>(n_jobs));
    std::vector<std::vector<double>> li(n, std::vector<double>(n_jobs)), em(n, std::vector<double>(n_jobs));
    std::vector<std::vector<int32_t>> cv(n, std::vector<int32_t>(n_jobs));
    std::vector<std::vector<uint64_t>> ol(n, std::vector<uint64_t>(n_jobs, 0));
    rc = multi_run(ctx, [&](int k, lpb200_ctx *sc) {
        const int g0 = ctx->bounds[k], gs = ctx->bounds[k + 1] - g0;
        std::vector<lpb200_params> ps(n_jobs);
        for (int j = 0; j < n_jobs; j++) {
            slice_params(ctx, M[j], &params[j], k, sp[k][j], true);
            ps[j] = sp[k][j].p;
        }
        std::vector<double> ll_k((size_t)n_jobs * gs);
        std::vector<int32_t> conv_k((size_t)n_jobs * gs);
        int r = lpb200_fit_model_multi(sc, n_jobs, models, ps.data(), c, mrm, li[k].data(), em[k].data(), cv[k].data(), ll_k.data(),
                                       conv_k.data(), ol[k].data());
        if (r) return r;
        for (int j = 0; j < n_jobs; j++) {
            memcpy(ll + (size_t)j * G + g0, ll_k.data() + (size_t)j * gs, sizeof(double) * gs);
            memcpy(conv + (size_t)j * G + g0, conv_k.data() + (size_t)j * gs, sizeof(int32_t) * gs);
        }
        return 0;
    });
    if (rc) return rc;
    for (int j = 0; j < n_jobs; j++) {
        for (int k = 0; k < n; k++) unslice_params(ctx, M[j], &params[j], k, sp[k][j], false);
        ll_init[j] = li[0][j];
        em_ll[j] = em[0][j];
        converged[j] = cv[0][j];
        if (ole) {
            ole[j] = 0;
            for (int k = 0; k < n; k++) ole[j] += ol[k][j];
        }
    }
    return 0;
}

static int multi_expand_fits(lpb200_ctx *ctx, const lpb200_model *m, const lpb200_params *p, const int32_t *gene_ids, int n_ids,
                             int what, double *z) {
    DevModel M;
    int rc = multi_model(ctx, m, M);
    if (rc) return rc;
    if (n_ids <= 0) return 0;
    if (!p || !gene_ids || !z) return set_err(ctx, LPB200_E_ARG, "expand_fits: NULL argument");
    const int G = ctx->G, N = ctx->N;
    for (int r = 0; r < n_ids; r++)
        if (gene_ids[r] < 0 || gene_ids[r] >= G) return set_err(ctx, LPB200_E_ARG, "gene id out of range");
    return multi_run(ctx, [&](int k, lpb200_ctx *sc) {
        const int g0 = ctx->bounds[k], g1 = ctx->bounds[k + 1];
        std::vector<int32_t> local, rows;
        for (int r = 0; r < n_ids; r++)
            if (gene_ids[r] >= g0 && gene_ids[r] < g1) {
                local.push_back(gene_ids[r] - g0);
                rows.push_back(r);
            }
        if (local.empty()) return 0;
        ShardParams sp;
        slice_params(ctx, M, p, k, sp, true);
        const int nl = (int)local.size();
        std::vector<double> zk((size_t)nl * N);
        int r = lpb200_expand_fits(sc, m, &sp.p, local.data(), nl, what, zk.data());
        if (r) return r;
        for (int j = 0; j < N; j++)
            for (int q = 0; q < nl; q++) z[rows[q] + (size_t)j * n_ids] = zk[q + (size_t)j * nl];
        return 0;
    });
}

static int multi_get_stats(lpb200_ctx *ctx, lpb200_stats *s) {
    *s = lpb200_stats{};
    for (lpb200_ctx *c : ctx->shards) {
        s->fn_evals += c->stats.fn_evals;
        s->ole += c->stats.ole;
        s->launches += c->stats.launches;
        s->kernel_ms = std::max(s->kernel_ms, c->stats.kernel_ms);
    }
    return 0;
}
