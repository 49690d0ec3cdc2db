// mirror_harness.cpp -- TEST INFRASTRUCTURE: a host mirror of fit_mudisp_kernel (lineagepulse_b200/csrc/lp_kernels.cuh).
//
// What is shared with the kernel (same source text, compiled here by g++ under LP_LIBM_MIRROR):
//   lp_eval.cuh    per-observation arithmetic of the specialised evaluator, table rule, request points
//   lp_model.cuh   theta -> natural scale, lp_cell_ll of the generic evaluator
//   lp_math.cuh    dnbinom (stirlerr / bd0), ZINB / NB terms, own log1p / lgamma
//   lp_libm.cuh    the DEVICE exp / log routines run through the host-emulation switch; the hardware
//                  reciprocal seed MUFU.RCP64H comes from the table dumped on a B200 (tests/golden/rcp64h.npz)
//   lp_bfgs.cuh    the resumable vmmin
// What is restated here: the kernel's evaluation layout -- lane-per-point / cells-per-warp slots, the
// stride of the cell loop over the gene's (non-zero first) cell permutation, the xor-shuffle reduction inside a
// (virtual) warp and the index-order reduction across the LP_VW virtual warps -- i.e. the ORDER in which the same
// doubles are added (independent of the CTA size and of the cluster size the kernel is launched with).
// tests/test_gpu_parity.py::test_fit_kernel_equals_host_mirror_bitwise asserts that kernel and mirror agree
// bit for bit on every (gene, start) item.  Not part of the product; nothing in lineagepulse_b200/ links it.
#define LP_LIBM_MIRROR 1
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "../../lineagepulse_b200/csrc/lp_eval.cuh"

extern "C" void mh_set_rcp64h_table(const int32_t *table) { lp_rcp64h_table = table; }

static void make_dev(const lpb200_design *d, const lpb200_model *m, DevDesign &D, DevModel &M) {
    D = DevDesign{};
    M = DevModel{};
    D.G = d->n_genes; D.N = d->n_cells; D.t = d->pseudotime; D.group = d->group_idx;
    D.sf = d->size_factors;
    if (D.sf) {   // lpb200_set_design drops size factors that are all 1
        bool all_one = true;
        for (int j = 0; j < D.N; j++) if (D.sf[j] != 1.0) { all_one = false; break; }
        if (all_one) D.sf = nullptr;
    }
    D.n_groups = d->group_idx ? d->n_groups : 0;
    if (d->spline_basis_mu && d->k_spline_mu > 0) { D.bs_mu = d->spline_basis_mu; D.k_mu = d->k_spline_mu; }
    if (d->spline_basis_disp && d->k_spline_disp > 0) { D.bs_disp = d->spline_basis_disp; D.k_disp = d->k_spline_disp; }
    D.n_conf_mu = d->n_conf_mu; D.bidx_mu = d->batch_idx_mu; D.n_conf_disp = d->n_conf_disp; D.bidx_disp = d->batch_idx_disp;
    for (int c = 0; c < LPB200_MAX_CONF; c++) {
        D.nb_mu[c] = c < d->n_conf_mu ? d->n_batches_mu[c] : 0;
        D.nb_disp[c] = c < d->n_conf_disp ? d->n_batches_disp[c] : 0;
        M.nbf_mu += D.nb_mu[c];
        M.nbf_disp += D.nb_disp[c];
    }
    if (d->pi_const_pred && d->n_pi_pred > 0) { D.pi_pred = d->pi_const_pred; D.P = d->n_pi_pred; }
    M.mu_model = m->mu_model; M.disp_model = m->disp_model; M.drop_model = m->drop_model;
    M.p_mu = lp_n_mu_params(m->mu_model, D.k_mu, D.n_groups);
    M.p_disp = lp_n_disp_params(m->disp_model, D.k_disp, D.n_groups);
    M.q_drop = m->drop_model == LPB200_DROP_NONE ? 0 : 1 + (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) + D.P;
    M.n_theta = M.p_disp + (M.nbf_disp - D.n_conf_disp) + M.p_mu + (M.nbf_mu - D.n_conf_mu);
}

struct GeneView {          // what CountView gives the kernel for one gene
    const int32_t *row;
    std::vector<uint32_t> perm;   // non-zero cells first (in cell order), then zeros; NA dropped (build_perm_kernel)
    int nnz = 0, nobs = 0, xmax = 0;
};

static void make_gene(const int32_t *row, int N, GeneView &g) {
    g.row = row;
    g.perm.clear();
    for (int j = 0; j < N; j++) if (row[j] > 0) { g.perm.push_back((uint32_t)j); g.xmax = std::max(g.xmax, row[j]); }
    g.nnz = (int)g.perm.size();
    for (int j = 0; j < N; j++) if (row[j] == 0) g.perm.push_back((uint32_t)j);
    g.nobs = (int)g.perm.size();
}

struct Shared {            // FitShared of the kernel
    PointNat pts[LP_KC];
    double vals[2 * LPB200_MAX_PARAMS];
    std::vector<double> tab[LP_NTAB];
    double tab_phi[LP_NTAB];
    int tab_of_point[LP_KC];
};

struct Ctx {
    DevDesign D;
    DevModel M;
    const double *drop;
    int use_table, use_fast;
};

// MU = -1: lp_eval_points (generic); otherwise lp_eval_points_fast<MU, ZINB>
template <int MU, bool ZINB>
static void eval_points(const Ctx &C, Shared &S, const GeneView &g, int gene, int nk, int ntabs, double *out_vals,
                        const double *pi_ptr, const double *l1m_ptr) {
    const DevDesign &D = C.D;
    const DevModel &M = C.M;
    const int N = D.N;
    const bool scalar_phi = lp_disp_is_scalar(D, M);
    const int tab_len = (MU >= 0 || scalar_phi) ? lp_table_len(C.use_table, ntabs, g.xmax, nk, N) : 0;
    if (tab_len > 0)
        for (int tsel = 0; tsel < ntabs; tsel++)
            for (int x = 0; x < tab_len; x++) S.tab[tsel][x] = (x == 0) ? 0.0 : lp_nb_xphi_part((double)x, S.tab_phi[tsel]);
    int npad = 1;
    while (npad < nk) npad <<= 1;
    const int cpw = 32 / npad, stride = LP_VW * cpw;
    // virtual warp v, slot s: cells r = v cpw + s + i LP_VW cpw of the gene's permutation, added in that order;
    // which physical warp / CTA of a cluster runs v is immaterial (lp_kernels.cuh)
    double red[LP_VW][LP_KC];
    for (int v = 0; v < LP_VW; v++) {
        double acc[32];
        for (int lane = 0; lane < 32; lane++) {
            const int k = lane & (npad - 1), slot = lane / npad;
            const bool active = k < nk;
            double a = 0.0;
            if (active) {
                const PointNat &P = S.pts[k];
                const double *tabp = tab_len > 0 ? S.tab[S.tab_of_point[k]].data() : nullptr;
                if constexpr (MU >= 0) {
                    FastPoint F;
                    lp_fast_point<MU>(P, F);
                    for (int r = v * cpw + slot; r < g.nobs; r += stride) {
                        const int j = (int)g.perm[r];
                        const int x = (r < g.nnz) ? g.row[j] : 0;
                        a += lp_fast_cell<MU, ZINB>(F, x, j, D.t, D.sf, pi_ptr, l1m_ptr, tabp, tab_len);
                    }
                } else {
                    double phi_s = 1.0, logphi_s = 0.0;
                    if (scalar_phi) {
                        phi_s = P.disp[0];
                        logphi_s = lp_log(phi_s);
                    }
                    for (int r = v * cpw + slot; r < g.nobs; r += stride) {
                        const int j = (int)g.perm[r];
                        const int x = g.row[j];
                        const double s = D.sf ? D.sf[j] : 1.0;
                        double pf = 0.0, l1 = 0.0;
                        if (pi_ptr) {
                            pf = pi_ptr[j];
                            l1 = l1m_ptr[j];
                        }
                        a += lp_cell_ll(D, M, P, C.drop, gene, j, x, s, pf, l1, phi_s, logphi_s, scalar_phi ? tabp : nullptr, tab_len);
                    }
                }
            }
            acc[lane] = a;
        }
        // acc += __shfl_xor_sync(acc, o) for o = 16 ... npad: every lane adds its partner's value of the previous step
        for (int o = 16; o >= npad; o >>= 1) {
            double nxt[32];
            for (int lane = 0; lane < 32; lane++) nxt[lane] = acc[lane] + acc[lane ^ o];
            memcpy(acc, nxt, sizeof acc);
        }
        for (int lane = 0; lane < npad && lane < nk; lane++) red[v][lane] = acc[lane];
    }
    for (int k = 0; k < nk; k++) {
        double s = 0.0;
        for (int v = 0; v < LP_VW; v++) s += red[v][k];
        out_vals[k] = s;
    }
}

template <int MU, bool ZINB>
static void fit_item(const Ctx &C, const GeneView &g, int gene, const double *theta0, int maxit, double reltol, double ndeps,
                     const double *pi_ptr, const double *l1m_ptr, double *theta_out, double *ll, int *conv, unsigned *evals) {
    const DevDesign &D = C.D;
    const DevModel &M = C.M;
    const int n = M.n_theta;
    const bool scalar_phi = lp_disp_is_scalar(D, M);
    Shared S;
    for (auto &t : S.tab) t.assign(LP_TAB, 0.0);
    BfgsState<LPB200_MAX_PARAMS> st;
    bfgs_start(st, n, theta0, maxit, reltol, ndeps);
    for (;;) {
        const int req = st.req;
        if (req == LP_REQ_NONE) break;
        const int total = (req == LP_REQ_F) ? 1 : 2 * n;
        for (int base = 0; base < total; base += LP_KC) {
            const int kc = std::min(LP_KC, total - base);
            for (int tid = 0; tid < kc; tid++) {
                const int tsel = lp_request_point(D, M, st.b, n, req, base + tid, ndeps, scalar_phi, S.pts[tid]);
                S.tab_of_point[tid] = tsel;
                if (scalar_phi) S.tab_phi[tsel] = S.pts[tid].disp[0];
            }
            if (scalar_phi && req == LP_REQ_GRAD) S.tab_phi[0] = lp_base_phi(st.b);
            eval_points<MU, ZINB>(C, S, g, gene, kc, lp_request_ntabs(req, base), S.vals + base, pi_ptr, l1m_ptr);
        }
        bfgs_advance(st, S.vals);
    }
    for (int i = 0; i < n; i++) theta_out[i] = st.b[i];
    *ll = (st.fail == 1001) ? NAN : -st.Fmin;
    *conv = st.fail;
    *evals = st.evals;
}

// X: G x N gene-major int32 (NA = -1); theta0 / theta_out: [n_items][n_theta] with item = gene * n_starts + start;
// drop: N x q column-major (or NULL); use_fast: take the specialised evaluator when the kernel launcher would.
extern "C" int mh_fit_items(const lpb200_design *d, const lpb200_model *m, const int32_t *X, const double *theta0, int n_starts,
                            const double *drop, int maxit, double reltol, double ndeps, int use_fast, int use_table, int nthreads,
                            double *theta_out, double *ll, int *conv, unsigned *evals) {
    Ctx C;
    make_dev(d, m, C.D, C.M);
    C.drop = drop;
    C.use_table = use_table;
    const DevDesign &D = C.D;
    const DevModel &M = C.M;
    const int G = D.G, N = D.N, n = M.n_theta;
    const bool fast = use_fast && lp_disp_is_scalar(D, M) && D.n_conf_mu == 0 &&
                      (M.mu_model == LPB200_MU_CONSTANT || M.mu_model == LPB200_MU_IMPULSE) &&
                      (M.drop_model == LPB200_DROP_NONE || M.drop_model == LPB200_DROP_LOGISTIC);
    C.use_fast = fast;
    // pi_vector_kernel: the gene-independent logistic drop-out rates and log(1 - pi)
    std::vector<double> pi_glob, l1m_glob;
    if (M.drop_model == LPB200_DROP_LOGISTIC && D.P == 0) {
        pi_glob.resize(N);
        l1m_glob.resize(N);
        for (int j = 0; j < N; j++) {
            pi_glob[j] = lp_dropout_rate(lp_drop_lin(D, M, drop, 0, j, 1.0));
            l1m_glob[j] = lp_log(1.0 - pi_glob[j]);
        }
    }
    std::atomic<int> next(0);
    const int n_items = G * n_starts;
    auto worker = [&]() {
        std::vector<double> pi_g, l1m_g;
        GeneView g;
        int have_gene = -1;
        for (;;) {
            const int item = next.fetch_add(1);
            if (item >= n_items) break;
            const int gene = item / n_starts;
            if (gene != have_gene) {
                g = GeneView();
                make_gene(X + (size_t)gene * N, N, g);
                have_gene = gene;
            }
            const double *pi_ptr = nullptr, *l1m_ptr = nullptr;
            if (M.drop_model == LPB200_DROP_LOGISTIC) {   // lp_fixed_pi
                if (D.P == 0) {
                    pi_ptr = pi_glob.data();
                    l1m_ptr = l1m_glob.data();
                } else {
                    pi_g.resize(N);
                    l1m_g.resize(N);
                    for (int j = 0; j < N; j++) {
                        pi_g[j] = lp_dropout_rate(lp_drop_lin(D, M, drop, gene, j, 1.0));
                        l1m_g[j] = lp_log(1.0 - pi_g[j]);
                    }
                    pi_ptr = pi_g.data();
                    l1m_ptr = l1m_g.data();
                }
            }
            const double *th0 = theta0 + (size_t)item * n;
            double *tho = theta_out + (size_t)item * n;
#define MH_FIT(MU, ZINB) fit_item<MU, ZINB>(C, g, gene, th0, maxit, reltol, ndeps, pi_ptr, l1m_ptr, tho, ll + item, conv + item, evals + item)
            if (!fast) MH_FIT(-1, false);
            else {
                const bool zinb = M.drop_model == LPB200_DROP_LOGISTIC;
                if (M.mu_model == LPB200_MU_IMPULSE) { if (zinb) MH_FIT(LP_MU_FAST_IMPULSE, true); else MH_FIT(LP_MU_FAST_IMPULSE, false); }
                else { if (zinb) MH_FIT(LP_MU_FAST_CONST, true); else MH_FIT(LP_MU_FAST_CONST, false); }
            }
#undef MH_FIT
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, nthreads); t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return n;
}

// the device exp / log / division routines on the host, for the bit-for-bit comparison with the GPU
extern "C" void mh_exp_log(const double *x, int n, double *e, double *l) {
    for (int k = 0; k < n; k++) {
        e[k] = lp_exp(x[k]);
        l[k] = lp_log(x[k]);
    }
}
extern "C" double mh_ll_obs(double x, double mu, double phi, double pi) {
    if (x == 0) return pi >= 0 ? lp_ll_zero_zinb(mu, phi, pi) : lp_ll_zero_nb(mu, phi, lp_log(phi));
    return lp_ll_nonzero(x, lp_nb_xphi_part(x, phi), mu, phi, pi >= 0 ? lp_log(1.0 - pi) : 0.0);
}
