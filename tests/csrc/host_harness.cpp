// host_harness.cpp -- TEST INFRASTRUCTURE: compiles the LP_HD headers of the CUDA library
// (lp_math.cuh, lp_model.cuh, lp_bfgs.cuh) with g++ and drives them sequentially on the CPU, so
// the scalar logic the kernels run (resumable vmmin, parameter transforms, per-observation terms)
// can be checked against the oracle on a box without a GPU.  It is NOT a fallback: the product
// library (liblpb200.so) never contains or calls this.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../lineagepulse_b200/csrc/lp_bfgs.cuh"
#include "../../lineagepulse_b200/csrc/lp_model.cuh"
#include "../../lineagepulse_b200/csrc/lp_exact.cuh"

static void make_dev(const lpb200_design *d, const lpb200_model *m, DevDesign &D, DevModel &M) {
    D.G = d->n_genes; D.N = d->n_cells; D.sf = d->size_factors; D.t = d->pseudotime; D.group = d->group_idx;
    D.n_groups = d->n_groups; D.bs_mu = d->spline_basis_mu; D.k_mu = d->k_spline_mu;
    D.bs_disp = d->spline_basis_disp; D.k_disp = d->k_spline_disp;
    D.n_conf_mu = d->n_conf_mu; D.bidx_mu = d->batch_idx_mu; D.n_conf_disp = d->n_conf_disp; D.bidx_disp = d->batch_idx_disp;
    M.nbf_mu = M.nbf_disp = 0;
    for (int c = 0; c < LPB200_MAX_CONF; c++) {
        D.nb_mu[c] = c < d->n_conf_mu ? d->n_batches_mu[c] : 0; D.nb_disp[c] = c < d->n_conf_disp ? d->n_batches_disp[c] : 0;
        M.nbf_mu += D.nb_mu[c]; M.nbf_disp += D.nb_disp[c];
    }
    D.pi_pred = d->pi_const_pred; D.P = d->n_pi_pred;
    M.mu_model = m->mu_model; M.disp_model = m->disp_model; M.drop_model = m->drop_model;
    M.p_mu = lp_n_mu_params(m->mu_model, D.k_mu, D.n_groups); M.p_disp = lp_n_disp_params(m->disp_model, D.k_disp, D.n_groups);
    M.q_drop = m->drop_model == LPB200_DROP_NONE ? 0 : 1 + (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) + D.P;
    M.n_theta = M.p_disp + (M.nbf_disp - D.n_conf_disp) + M.p_mu + (M.nbf_mu - D.n_conf_mu);
}

extern "C" double lporacle_ll_zinb_obs(double x, double mu, double disp, double pi);
static int g_use_oracle_obs = 0;
static double *g_trace = nullptr; static int g_ntrace = 0;
extern "C" void hh_set_trace(double *t) { g_trace = t; g_ntrace = 0; }
extern "C" int hh_ntrace() { return g_ntrace; }
extern "C" void hh_use_oracle_obs(int v) { g_use_oracle_obs = v; }
static double eval_theta(const DevDesign &D, const DevModel &M, const int32_t *x, const double *th, const double *drop,
                         int gene, const double *pi_fixed) {
    PointNat g;
    lp_unpack_theta(D, M, th, g);
    double phi = g.disp[0], logphi = log(phi);
    if (g_use_oracle_obs) {
        long double sz = 0, snz = 0;
        for (int j = 0; j < D.N; j++) {
            if (x[j] < 0) continue;
            double mu = lp_mu_at(D, M, g, j) * (D.sf ? D.sf[j] : 1.0);
            double v = lporacle_ll_zinb_obs((double)x[j], mu, lp_disp_at(D, M, g, j), pi_fixed ? pi_fixed[j] : -1.0);
            if (x[j] == 0) sz += v; else snz += v;
        }
        return (double)sz + (double)snz;
    }
    double acc = 0.0;
    for (int j = 0; j < D.N; j++) {
        double pf = pi_fixed ? pi_fixed[j] : 0.0;
        double l1 = pi_fixed ? log(1.0 - pf) : 0.0;
        acc += lp_cell_ll(D, M, g, drop, gene, j, x[j],
                          D.sf ? D.sf[j] : 1.0, pf, l1, phi, logphi, nullptr, 0);
    }
    return acc;
}

extern "C" int hh_fit_gene(const lpb200_design *d, const lpb200_model *m, const int32_t *x, int gene, const double *theta0,
                           const double *drop, const double *pi_fixed, int maxit, double reltol, double ndeps,
                           double *theta_out, double *ll, int *conv, unsigned *evals) {
    DevDesign D; DevModel M;
    make_dev(d, m, D, M);
    BfgsState<LPB200_MAX_PARAMS> st;
    bfgs_start(st, M.n_theta, theta0, maxit, reltol, ndeps);
    double vals[2 * LPB200_MAX_PARAMS], th[LPB200_MAX_PARAMS];
    while (st.req != LP_REQ_NONE) {
        if (st.req == LP_REQ_F) {
            vals[0] = eval_theta(D, M, x, st.b, drop, gene, pi_fixed);
        } else {
            for (int k = 0; k < 2 * st.n; k++) {
                memcpy(th, st.b, sizeof(double) * st.n);
                int c = bfgs_grad_coord(k);
                th[c] = bfgs_grad_point(st.b[c], k, st.ndeps);
                vals[k] = eval_theta(D, M, x, th, drop, gene, pi_fixed);
            }
        }
        bfgs_advance(st, vals);
        if (g_trace && g_ntrace < 4096) { g_trace[g_ntrace * 3] = st.f; g_trace[g_ntrace * 3 + 1] = st.Fmin; g_trace[g_ntrace * 3 + 2] = st.phase * 1000 + st.iter; g_ntrace++; }
    }
    memcpy(theta_out, st.b, sizeof(double) * st.n);
    *ll = -st.Fmin; *conv = st.fail; *evals = st.evals;
    return M.n_theta;
}

static double rosen(const double *x) { return 100 * (x[1] - x[0] * x[0]) * (x[1] - x[0] * x[0]) + (1 - x[0]) * (1 - x[0]); }
extern "C" int hh_rosenbrock(int maxit, double reltol, double *par, double *value, int *fncount, int *grcount, int *conv) {
    BfgsState<4> st;
    double th0[2] = {-1.2, 1.0}, vals[4], th[2];
    bfgs_start(st, 2, th0, maxit, reltol, 1e-3);
    while (st.req != LP_REQ_NONE) {
        if (st.req == LP_REQ_F) vals[0] = -rosen(st.b);
        else for (int k = 0; k < 4; k++) {
            th[0] = st.b[0]; th[1] = st.b[1];
            th[bfgs_grad_coord(k)] = bfgs_grad_point(st.b[bfgs_grad_coord(k)], k, st.ndeps);
            vals[k] = -rosen(th);
        }
        bfgs_advance(st, vals);
    }
    par[0] = st.b[0]; par[1] = st.b[1]; *value = st.Fmin; *fncount = st.funcount; *grcount = st.gradcount; *conv = st.fail;
    return 0;
}

extern "C" double hh_ll_obs(double x, double mu, double phi, double pi) {
    double logphi = log(phi);
    if (x == 0) return pi >= 0 ? lp_ll_zero_zinb(mu, phi, pi) : lp_ll_zero_nb(mu, phi, logphi);
    if (lp_nb_small_x(x, phi)) {   // as lp_cell_ll: dnbinom_mu's x < 1e-10 * size branch
        double v = (pi >= 0 ? log(1.0 - pi) : 0.0) + lp_dnbinom_small_x(x, mu, phi);
        return v < LP_LL_FLOOR ? LP_LL_FLOOR : v;
    }
    return lp_ll_nonzero(x, lp_nb_xphi_part(x, phi), mu, phi, pi >= 0 ? log(1.0 - pi) : 0.0);
}

// exact fixed-point accumulation (lp_exact.cuh): result words for the given summation order
extern "C" double hh_exact_sum(const double *v, int n, long long *words) {
    ExactSum s = {0, 0, 0};
    for (int i = 0; i < n; i++) lp_exact_add(s, v[i]);
    words[0] = s.a; words[1] = s.b; words[2] = s.bad;
    return lp_exact_value(s.a, s.b, s.bad);
}
