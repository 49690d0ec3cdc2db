/*
 * lp_oracle.c -- CPU oracle for the LineagePulse model-fitting hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see lp_oracle.h).  A plain-C restatement of the
 * reference algorithm, one function per reference function, each citing the
 * reference file:line it follows (paths relative to /root/reference/R/, prefix
 * srcLineagePulse_ dropped).  Arithmetic that the reference delegates to R base
 * (not vendored under /root/reference, no pinned R version; package dated
 * 2017-10 => R 3.4 era) is restated from the published algorithms:
 *   - stats::optim(method="BFGS")  = vmmin, J.C. Nash, Compact Numerical Methods
 *     for Computers, Alg. 21, with optim's central-difference gradient (ndeps);
 *   - stats::dnbinom(mu=,size=,log=TRUE) = nmath dnbinom_mu -> dbinom_raw with
 *     stirlerr/bd0, C. Loader, "Fast and Accurate Computation of Binomial
 *     Probabilities" (2000);
 *   - stats::lm(y ~ 0 + B) = least squares by QR;
 *   - stats::pchisq(lower.tail=FALSE), stats::p.adjust(method="BH");
 *   - base::sum accumulates in long double.
 * PARITY STATUS: the reference ships no golden vectors and R is not installed, so
 * parity is unpinned by an R session.  Pinned instead against (1) the reference's own
 * R source executed by tests/r_eval.py (fitModel and everything under it, unmodified;
 * tests/golden/reference_exec.json: reproduced to the last bit, tests/test_reference_exec.py)
 * with only R's compiled builtins restated, and (2) for those builtins: mpmath and scipy
 * golden values, the ?optim Rosenbrock known answers and analytic identities in tests/.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lp_oracle.h"

#define LL_FLOOR (-700.0)              /* scaLogLikPrecLim, evalLogLik.R:37,123 */
#define LN_SQRT_2PI 0.918938533204672741780329736406L
#define LN_2PI 1.837877066409345483560659472811
#define MAXB 32                        /* batches per confounder */

/* ------------------------------------------------------------------------- */
/* nmath restatement: stirlerr, bd0, dbinom_raw, dnbinom_mu (log scale only)  */
/* ------------------------------------------------------------------------- */

static double sferr_halves[31];

/* stirlerr(n) = log(n!) - log( sqrt(2*pi*n)*(n/e)^n ); table for n = 0, 0.5, ..., 15 */
static void __attribute__((constructor)) init_sferr(void) {
    sferr_halves[0] = 0.0;
    for (int k = 1; k <= 30; k++) {
        long double n = 0.5L * k;
        sferr_halves[k] = (double)(lgammal(n + 1.0L) - (n + 0.5L) * logl(n) + n - LN_SQRT_2PI);
    }
}

static double stirlerr(double n) {
    const double S0 = 0.083333333333333333333;        /* 1/12 */
    const double S1 = 0.00277777777777777777778;      /* 1/360 */
    const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
    const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
    const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
    double nn;
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (int)nn) return sferr_halves[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - (double)LN_SQRT_2PI;
    }
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* bd0(x, np) = x log(x/np) + np - x, stable when x ~ np */
static double bd0(double x, double np) {
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

static double dbinom_raw_log(double x, double n, double p, double q) {
    double lf, lc;
    if (p == 0) return (x == 0) ? 0.0 : -INFINITY;
    if (q == 0) return (x == n) ? 0.0 : -INFINITY;
    if (x == 0) {
        if (n == 0) return 0.0;
        lc = (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
        return lc;
    }
    if (x == n) {
        lc = (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
        return lc;
    }
    if (x < 0 || x > n) return -INFINITY;
    lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    lf = LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

/* dnbinom(x, mu=mu, size=size, log=TRUE) */
double lporacle_dnbinom_mu_log(double x, double size, double mu) {
    if (isnan(x) || isnan(size) || isnan(mu)) return x + size + mu;
    if (mu < 0 || size < 0) return NAN;
    if (x < 0 || !isfinite(x)) return -INFINITY;
    if (x == 0 && size == 0) return 0.0;
    if (x == 0)
        return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    if (x < 1e-10 * size) {
        double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
        return x * p - mu - lgamma(x + 1) + log1p(x * (x - 1) / (2 * size));
    } else {
        double p = size / (size + x);
        double ans = dbinom_raw_log(size, x + size, size / (size + mu), mu / (size + mu));
        return log(p) + ans;
    }
}

/* pchisq(q, df, lower.tail=FALSE) for integer df (df is a difference of parameter counts,
 * runHypothesisTests.R:75-83): closed forms of Q(df/2, q/2). */
double lporacle_pchisq_upper(double q, double df) {
    if (isnan(q) || isnan(df)) return NAN;
    if (df < 0) return NAN;
    if (q <= 0) return 1.0;
    if (df == 0) return 0.0;
    if (!isfinite(q)) return 0.0;
    long double x = 0.5L * q;
    int idf = (int)df;
    if ((double)idf != df) return NAN; /* oracle covers integer df only */
    if (idf % 2 == 0) {
        int m = idf / 2; /* Q = exp(-x) sum_{k<m} x^k/k! */
        long double s = 0.0L, term = 1.0L;
        for (int j = 0; j < m; j++) { /* normalised by the k = m-1 term */
            s += term;
            term *= (long double)(m - 1 - j) / x;
        }
        long double lq = -x + (m - 1) * logl(x) - lgammal((long double)m) + logl(s);
        return (double)expl(lq);
    } else {
        int m = (idf - 1) / 2; /* Q(m+1/2,x) = erfc(sqrt x) + exp(-x) sum_{k<m} x^(k+1/2)/Gamma(k+3/2) */
        long double res = erfcl(sqrtl(x));
        for (int k = 0; k < m; k++)
            res += expl(-x + (k + 0.5L) * logl(x) - lgammal(k + 1.5L));
        return (double)res;
    }
}

/* ------------------------------------------------------------------------- */
/* scalar model pieces                                                        */
/* ------------------------------------------------------------------------- */

/* evalImpulseModel, evalImpulseModel.R:23-37; par = (beta1,beta2,h0,h1,h2,t1,t2) */
double lporacle_impulse_value(const double *p, double t) {
    double t1 = t - p[5];
    double t2 = t - p[6];
    return (1 / p[3]) * (p[2] + (p[3] - p[2]) * (1 / (1 + exp(-p[0] * t1)))) *
           (p[4] + (p[3] - p[4]) * (1 / (1 + exp(p[1] * t2))));
}

/* evalDropoutModel, evalDropoutModel.R:24-32 */
double lporacle_dropout_rate(const double *theta, const double *pred, int q) {
    const double off = 0.001;
    double lin = 0.0;
    for (int k = 0; k < q; k++) lin += pred[k] * theta[k];
    return off + (1 - 2 * off) * 1 / (1 + exp(-lin));
}

/* one observation of evalLogLikZINB (evalLogLik.R:107-149) / evalLogLikNB (:30-53);
 * mu already includes the size factor; pi < 0 selects NB. */
double lporacle_ll_zinb_obs(double x, double mu, double disp, double pi) {
    double v;
    if (pi >= 0) {
        if (x == 0)
            v = log((1 - pi) * pow(disp / (disp + mu), disp) + pi);
        else
            v = log(1 - pi) + lporacle_dnbinom_mu_log(x, disp, mu);
    } else {
        if (x == 0)
            v = disp * (log(disp) - log(disp + mu));
        else
            v = lporacle_dnbinom_mu_log(x, disp, mu);
    }
    if (v < LL_FLOOR) v = LL_FLOOR;
    return v;
}

/* ------------------------------------------------------------------------- */
/* per-gene natural-scale parameters and their expansion (decompress*)        */
/* ------------------------------------------------------------------------- */

typedef struct {
    double mu[LPB200_MAX_PARAMS];
    double disp[LPB200_MAX_PARAMS];
    double bmu[LPB200_MAX_CONF][MAXB];   /* incl. leading 1 */
    double bdisp[LPB200_MAX_CONF][MAXB];
} gene_nat;

static int n_mu_par(const lpb200_model *m, const lpb200_design *d) {
    switch (m->mu_model) {
    case LPB200_MU_CONSTANT: return 1;
    case LPB200_MU_IMPULSE: return 7;
    case LPB200_MU_SPLINES: return d->k_spline_mu;
    case LPB200_MU_GROUPS: return d->n_groups;
    default: return -1;
    }
}
static int n_disp_par(const lpb200_model *m, const lpb200_design *d) {
    switch (m->disp_model) {
    case LPB200_DISP_CONSTANT: return 1;
    case LPB200_DISP_SPLINES: return d->k_spline_disp;
    case LPB200_DISP_GROUPS: return d->n_groups;
    default: return -1;
    }
}
static int n_drop_par(const lpb200_model *m, const lpb200_design *d) {
    if (m->drop_model == LPB200_DROP_NONE) return 0;
    return 1 + (m->drop_model == LPB200_DROP_LOGISTIC_OFMU ? 1 : 0) + d->n_pi_pred;
}
static int n_batch_cols(const int32_t *nb, int nconf) {
    int s = 0;
    for (int c = 0; c < nconf; c++) s += nb[c];
    return s;
}

/* decompressMeansByGene with vecInterval=j, decompressParameters.R:26-88 (mu WITHOUT size factor) */
static inline double mu_at(const lpb200_design *d, int mu_model, const gene_nat *g, int j) {
    double mu;
    switch (mu_model) {
    case LPB200_MU_CONSTANT: mu = g->mu[0]; break;
    case LPB200_MU_IMPULSE: mu = lporacle_impulse_value(g->mu, d->pseudotime[j]); break;
    case LPB200_MU_SPLINES: {
        double lin = 0.0;
        for (int k = 0; k < d->k_spline_mu; k++)
            lin += d->spline_basis_mu[j + (size_t)k * d->n_cells] * g->mu[k];
        mu = exp(lin);
        break;
    }
    case LPB200_MU_GROUPS: mu = g->mu[d->group_idx[j]]; break;
    default: mu = NAN;
    }
    for (int c = 0; c < d->n_conf_mu; c++)
        mu = mu * g->bmu[c][d->batch_idx_mu[(size_t)c * d->n_cells + j]];
    return mu;
}

/* decompressDispByGene, decompressParameters.R:155-211 */
static inline double disp_at(const lpb200_design *d, int disp_model, const gene_nat *g, int j) {
    double phi;
    switch (disp_model) {
    case LPB200_DISP_CONSTANT: phi = g->disp[0]; break;
    case LPB200_DISP_SPLINES: {
        double lin = 0.0;
        for (int k = 0; k < d->k_spline_disp; k++)
            lin += d->spline_basis_disp[j + (size_t)k * d->n_cells] * g->disp[k];
        phi = exp(lin);
        break;
    }
    case LPB200_DISP_GROUPS: phi = g->disp[d->group_idx[j]]; break;
    default: phi = NAN;
    }
    for (int c = 0; c < d->n_conf_disp; c++)
        phi = phi * g->bdisp[c][d->batch_idx_disp[(size_t)c * d->n_cells + j]];
    return phi;
}

/* row i of the stored (natural-scale) parameter matrices -> gene_nat */
static void load_gene_nat(const lpb200_design *d, const lpb200_model *m, const lpb200_params *p,
                          int i, gene_nat *g) {
    int G = d->n_genes;
    int pm = n_mu_par(m, d), pd = n_disp_par(m, d);
    for (int k = 0; k < pm; k++) g->mu[k] = p->mat_mu[i + (size_t)k * G];
    for (int k = 0; k < pd; k++) g->disp[k] = p->mat_disp[i + (size_t)k * G];
    size_t off = 0;
    for (int c = 0; c < d->n_conf_mu; c++) {
        for (int b = 0; b < d->n_batches_mu[c]; b++) g->bmu[c][b] = p->batch_mu[off + i + (size_t)b * G];
        off += (size_t)G * d->n_batches_mu[c];
    }
    off = 0;
    for (int c = 0; c < d->n_conf_disp; c++) {
        for (int b = 0; b < d->n_batches_disp[c]; b++) g->bdisp[c][b] = p->batch_disp[off + i + (size_t)b * G];
        off += (size_t)G * d->n_batches_disp[c];
    }
}

static void store_gene_nat(const lpb200_design *d, const lpb200_model *m, lpb200_params *p, int i,
                           const gene_nat *g) {
    int G = d->n_genes;
    int pm = n_mu_par(m, d), pd = n_disp_par(m, d);
    for (int k = 0; k < pm; k++) p->mat_mu[i + (size_t)k * G] = g->mu[k];
    for (int k = 0; k < pd; k++) p->mat_disp[i + (size_t)k * G] = g->disp[k];
    size_t off = 0;
    for (int c = 0; c < d->n_conf_mu; c++) {
        for (int b = 0; b < d->n_batches_mu[c]; b++) p->batch_mu[off + i + (size_t)b * G] = g->bmu[c][b];
        off += (size_t)G * d->n_batches_mu[c];
    }
    off = 0;
    for (int c = 0; c < d->n_conf_disp; c++) {
        for (int b = 0; b < d->n_batches_disp[c]; b++) p->batch_disp[off + i + (size_t)b * G] = g->bdisp[c][b];
        off += (size_t)G * d->n_batches_disp[c];
    }
}

/* drop-out rate of gene i at cell j (decompressDropoutRateByGene, decompressParameters.R:295-320);
 * mu is the mean WITHOUT size factor (fitMeanDispersion.R:204-208, evalLogLik.R:388-392) */
static inline double pi_at(const lpb200_design *d, const lpb200_model *m, const double *mat_drop,
                           int i, int j, double mu) {
    double pred[2 + LPB200_MAX_PI_PRED], theta[2 + LPB200_MAX_PI_PRED];
    int q = 0;
    pred[q++] = 1.0;
    if (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) pred[q++] = log(mu);
    for (int k = 0; k < d->n_pi_pred; k++) pred[q++] = d->pi_const_pred[i + (size_t)k * d->n_genes];
    for (int k = 0; k < q; k++) theta[k] = mat_drop[j + (size_t)k * d->n_cells];
    return lporacle_dropout_rate(theta, pred, q);
}

/* evalLogLikGene -> evalLogLikZINB / evalLogLikNB for one gene (evalLogLik.R:30-53,107-149,202-228):
 * zeros and non-zeros are summed separately (R sum = long double accumulation) */
/* Sensitivity knob (tests only): 0 = R on x86 (sum() accumulates in 80-bit long double),
 * 1 = R built without long double (plain double accumulation).  Both are legitimate R
 * behaviours; the spread between them is the reference algorithm's own numerical conditioning. */
static int g_sum_mode = 0;
void lporacle_set_sum_mode(int mode) { g_sum_mode = mode; }

static double ll_gene(const int32_t *x, int N, const double *mu, const double *sf, const double *disp,
                      const double *pi) {
    if (g_sum_mode == 1) {
        double dz = 0.0, dnz = 0.0;
        for (int j = 0; j < N; j++) {
            int32_t xj = x[j];
            if (xj < 0) continue;
            double mus = sf ? mu[j] * sf[j] : mu[j] * 1.0;
            double v = lporacle_ll_zinb_obs((double)xj, mus, disp[j], pi ? pi[j] : -1.0);
            if (xj == 0) dz += v; else dnz += v;
        }
        return dz + dnz;
    }
    long double sz = 0.0L, snz = 0.0L;
    for (int j = 0; j < N; j++) {
        int32_t xj = x[j];
        if (xj < 0) continue; /* NA: in neither vecidxZero nor vecidxNotZero */
        double mus = sf ? mu[j] * sf[j] : mu[j] * 1.0;
        double v = lporacle_ll_zinb_obs((double)xj, mus, disp[j], pi ? pi[j] : -1.0);
        if (xj == 0) sz += v; else snz += v;
    }
    return (double)sz + (double)snz;
}

/* ------------------------------------------------------------------------- */
/* thread pool over independent work items (the BiocParallel::bplapply analogue) */
/* ------------------------------------------------------------------------- */

typedef void (*item_fn)(int item, void *ctx, void *scratch);
typedef struct {
    int n_items;
    volatile int next;
    item_fn fn;
    void *ctx;
    size_t scratch_bytes;
} pool_t;

static void *pool_worker(void *arg) {
    pool_t *p = (pool_t *)arg;
    void *scratch = p->scratch_bytes ? malloc(p->scratch_bytes) : NULL;
    for (;;) {
        int it = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (it >= p->n_items) break;
        p->fn(it, p->ctx, scratch);
    }
    free(scratch);
    return NULL;
}

static void run_pool(int n_items, int nthreads, item_fn fn, void *ctx, size_t scratch_bytes) {
    pool_t p = {n_items, 0, fn, ctx, scratch_bytes};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n_items) nthreads = n_items > 0 ? n_items : 1;
    if (nthreads == 1) {
        pool_worker(&p);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pool_worker, &p);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------- */
/* evalLogLikMatrix (evalLogLik.R:307-431)                                     */
/* ------------------------------------------------------------------------- */

typedef struct {
    const int32_t *counts;
    const lpb200_design *d;
    const lpb200_model *m;
    const lpb200_params *p;
    double *ll;
} evalmat_ctx;

static double eval_gene_ll_stored(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                                  const lpb200_params *p, int i, double *buf /* 3N */) {
    int N = d->n_cells;
    double *mu = buf, *disp = buf + N, *pi = buf + 2 * (size_t)N;
    gene_nat g;
    load_gene_nat(d, m, p, i, &g);
    for (int j = 0; j < N; j++) {
        mu[j] = mu_at(d, m->mu_model, &g, j);
        disp[j] = disp_at(d, m->disp_model, &g, j);
    }
    int has_pi = m->drop_model != LPB200_DROP_NONE;
    if (has_pi)
        for (int j = 0; j < N; j++) pi[j] = pi_at(d, m, p->mat_drop, i, j, mu[j]);
    return ll_gene(counts + (size_t)i * N, N, mu, d->size_factors, disp, has_pi ? pi : NULL);
}

static void evalmat_item(int i, void *vctx, void *scratch) {
    evalmat_ctx *c = (evalmat_ctx *)vctx;
    c->ll[i] = eval_gene_ll_stored(c->counts, c->d, c->m, c->p, i, (double *)scratch);
}

int lporacle_eval_loglik(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         const lpb200_params *p, double *ll, int nthreads) {
    if (m->mu_model == LPB200_MU_MM) return LPB200_E_NOTSUP; /* evalLogLik.R:271 stop() */
    if (n_mu_par(m, d) < 0 || n_disp_par(m, d) < 0) return LPB200_E_ARG;
    evalmat_ctx c = {counts, d, m, p, ll};
    run_pool(d->n_genes, nthreads, evalmat_item, &c, sizeof(double) * 3 * (size_t)d->n_cells);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* stats::optim(method="BFGS"): vmmin + central-difference gradient            */
/* ------------------------------------------------------------------------- */

typedef double (*opt_fn)(int n, const double *par, void *ex);
typedef void (*opt_gr)(int n, const double *par, double *g, void *ex);

#define STEPREDN 0.2
#define ACCTOL 0.0001
#define RELTEST 10.0

typedef struct {
    opt_fn fn;       /* returns the value to MAXIMISE (fnscale = -1) */
    void *ex;
    double ndeps;
    uint64_t evals;  /* all objective evaluations, gradient ones included */
    int nonfinite;   /* R would have raised an error */
} optim_env;

/* fminfn of optim.c: val = fn(par)/fnscale with fnscale = -1, parscale = 1 */
static double fminfn(int n, const double *b, optim_env *e) {
    e->evals++;
    return e->fn(n, b, e->ex) / (-1.0);
}

/* fmingr of optim.c with gr = NULL: central differences, eps = ndeps */
static void fmingr_num(int n, double *b, double *g, optim_env *e) {
    for (int i = 0; i < n; i++) {
        double eps = e->ndeps;
        double save = b[i];
        b[i] = save + eps;
        double v1 = fminfn(n, b, e);
        b[i] = save - eps;
        double v2 = fminfn(n, b, e);
        g[i] = (v1 - v2) / (2 * eps);
        if (!isfinite(g[i])) e->nonfinite = 1; /* "non-finite finite-difference value [i]" */
        b[i] = save;
    }
}

/* vmmin (Nash Alg. 21 as in R's optim.c); gr == NULL => numerical gradient.
 * Returns 0, or LPB200_E_NONFINITE where R would error. */
static int vmmin(int n, double *b, double *Fmin, optim_env *e, opt_gr gr, int maxit, double abstol,
                 double reltol, int *fncount, int *grcount, int *fail) {
    int accpoint, enough;
    double g[LPB200_MAX_PARAMS], t[LPB200_MAX_PARAMS], X[LPB200_MAX_PARAMS], c[LPB200_MAX_PARAMS];
    double B[LPB200_MAX_PARAMS][LPB200_MAX_PARAMS];
    int count = 0, funcount, gradcount;
    double f, gradproj;
    int i, j, ilast, iter = 0;
    double s, steplength;
    double D1, D2;

    if (maxit <= 0) {
        *fail = 0;
        *Fmin = fminfn(n, b, e);
        *fncount = *grcount = 0;
        return 0;
    }
    f = fminfn(n, b, e);
    if (!isfinite(f)) return LPB200_E_NONFINITE; /* "initial value in 'vmmin' is not finite" */
    *Fmin = f;
    funcount = gradcount = 1;
    if (gr) gr(n, b, g, e->ex); else fmingr_num(n, b, g, e);
    if (e->nonfinite) return LPB200_E_NONFINITE;
    iter++;
    ilast = gradcount;

    do {
        if (ilast == gradcount) {
            for (i = 0; i < n; i++) {
                for (j = 0; j < i; j++) B[i][j] = 0.0;
                B[i][i] = 1.0;
            }
        }
        for (i = 0; i < n; i++) {
            X[i] = b[i];
            c[i] = g[i];
        }
        gradproj = 0.0;
        for (i = 0; i < n; i++) {
            s = 0.0;
            for (j = 0; j <= i; j++) s -= B[i][j] * c[j];
            for (j = i + 1; j < n; j++) s -= B[j][i] * c[j];
            t[i] = s;
            gradproj += s * c[i];
        }

        if (gradproj < 0.0) { /* search direction is downhill */
            steplength = 1.0;
            accpoint = 0;
            do {
                count = 0;
                for (i = 0; i < n; i++) {
                    b[i] = X[i] + steplength * t[i];
                    if (RELTEST + X[i] == RELTEST + b[i]) count++; /* no change */
                }
                if (count < n) {
                    f = fminfn(n, b, e);
                    funcount++;
                    accpoint = isfinite(f) && (f <= *Fmin + gradproj * steplength * ACCTOL);
                    if (!accpoint) steplength *= STEPREDN;
                }
            } while (!(count == n || accpoint));
            enough = (f > abstol) && fabs(f - *Fmin) > reltol * (fabs(*Fmin) + reltol);
            if (!enough) {
                count = n;
                *Fmin = f;
            }
            if (count < n) { /* making progress */
                *Fmin = f;
                if (gr) gr(n, b, g, e->ex); else fmingr_num(n, b, g, e);
                if (e->nonfinite) return LPB200_E_NONFINITE;
                gradcount++;
                iter++;
                D1 = 0.0;
                for (i = 0; i < n; i++) {
                    t[i] = steplength * t[i];
                    c[i] = g[i] - c[i];
                    D1 += t[i] * c[i];
                }
                if (D1 > 0) {
                    D2 = 0.0;
                    for (i = 0; i < n; i++) {
                        s = 0.0;
                        for (j = 0; j <= i; j++) s += B[i][j] * c[j];
                        for (j = i + 1; j < n; j++) s += B[j][i] * c[j];
                        X[i] = s;
                        D2 += s * c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (i = 0; i < n; i++)
                        for (j = 0; j <= i; j++)
                            B[i][j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
                } else { /* D1 < 0 */
                    ilast = gradcount;
                }
            } else { /* no progress */
                if (ilast < gradcount) {
                    count = 0;
                    ilast = gradcount;
                }
            }
        } else { /* uphill search */
            count = 0;
            if (ilast == gradcount) count = n;
            else ilast = gradcount;
        }
        if (iter >= maxit) break;
        if (gradcount - ilast > 2 * n) ilast = gradcount; /* periodic restart */
    } while (count != n || ilast != gradcount);

    *fail = (iter < maxit) ? 0 : 1;
    *fncount = funcount;
    *grcount = gradcount;
    return 0;
}

/* ?optim Rosenbrock example: known-answer test of the restatement */
static double rosen_neg(int n, const double *x, void *ex) {
    (void)n; (void)ex;
    double x1 = x[0], x2 = x[1];
    return -(100 * (x2 - x1 * x1) * (x2 - x1 * x1) + (1 - x1) * (1 - x1));
}
static void rosen_gr(int n, const double *x, double *g, void *ex) {
    (void)n; (void)ex;
    double x1 = x[0], x2 = x[1];
    g[0] = -400 * x1 * (x2 - x1 * x1) - 2 * (1 - x1);
    g[1] = 200 * (x2 - x1 * x1);
}
int lporacle_vmmin_rosenbrock(int use_numgrad, int maxit, double reltol, double *par, double *value,
                              int *fncount, int *grcount, int *conv) {
    optim_env e = {rosen_neg, NULL, 1e-3, 0, 0};
    par[0] = -1.2;
    par[1] = 1.0;
    double fmin;
    int rc = vmmin(2, par, &fmin, &e, use_numgrad ? NULL : rosen_gr, maxit, -INFINITY, reltol,
                   fncount, grcount, conv);
    *value = fmin; /* optim reports Fmin*fnscale of the minimised -fn = the Rosenbrock value */
    return rc;
}

/* ------------------------------------------------------------------------- */
/* evalLogLikMuDispGeneFit (fitMeanDispersion.R:62-222) and fitMuDispGene      */
/* ------------------------------------------------------------------------- */

static inline double clampd(double v, double lo, double hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

/* theta = [disp params | disp batch logs | mu params | mu batch logs] -> natural scale with the
 * reference clamps (fitMeanDispersion.R:78-128,131-200 == :391-498) */
static void unpack_theta(const lpb200_design *d, const lpb200_model *m, const double *th, gene_nat *g) {
    const double LO = pow(10.0, -10.0), HI = pow(10.0, 10.0);
    int u = 0;
    if (m->disp_model == LPB200_DISP_CONSTANT) {
        g->disp[0] = clampd(exp(th[u]), LO, HI);
        u += 1;
    } else if (m->disp_model == LPB200_DISP_GROUPS) {
        for (int k = 0; k < d->n_groups; k++) g->disp[k] = clampd(exp(th[u + k]), LO, HI);
        u += d->n_groups;
    } else { /* splines: raw coefficients */
        for (int k = 0; k < d->k_spline_disp; k++)
            g->disp[k] = clampd(th[u + k], -10 * log(10.0), 10 * log(10.0));
        u += d->k_spline_disp;
    }
    for (int c = 0; c < d->n_conf_disp; c++) {
        g->bdisp[c][0] = 1.0;
        for (int b = 1; b < d->n_batches_disp[c]; b++) g->bdisp[c][b] = clampd(exp(th[u + b - 1]), LO, HI);
        u += d->n_batches_disp[c] - 1;
    }
    if (m->mu_model == LPB200_MU_CONSTANT) {
        g->mu[0] = clampd(exp(th[u]), LO, HI);
        u += 1;
    } else if (m->mu_model == LPB200_MU_GROUPS) {
        for (int k = 0; k < d->n_groups; k++) g->mu[k] = clampd(exp(th[u + k]), LO, HI);
        u += d->n_groups;
    } else if (m->mu_model == LPB200_MU_SPLINES) {
        for (int k = 0; k < d->k_spline_mu; k++) g->mu[k] = clampd(th[u + k], -HI, HI);
        u += d->k_spline_mu;
    } else { /* impulse: slots 3:5 are logs; betas floored only; t1,t2 free */
        for (int k = 0; k < 7; k++) g->mu[k] = th[u + k];
        for (int k = 2; k < 5; k++) g->mu[k] = exp(g->mu[k]);
        for (int k = 0; k < 2; k++)
            if (g->mu[k] < LO) g->mu[k] = LO;
        for (int k = 2; k < 5; k++) g->mu[k] = clampd(g->mu[k], LO, HI);
        u += 7;
    }
    for (int c = 0; c < d->n_conf_mu; c++) {
        g->bmu[c][0] = 1.0;
        for (int b = 1; b < d->n_batches_mu[c]; b++) g->bmu[c][b] = clampd(exp(th[u + b - 1]), LO, HI);
        u += d->n_batches_mu[c] - 1;
    }
}

static int n_theta(const lpb200_design *d, const lpb200_model *m) {
    return n_disp_par(m, d) + (n_batch_cols(d->n_batches_disp, d->n_conf_disp) - d->n_conf_disp) +
           n_mu_par(m, d) + (n_batch_cols(d->n_batches_mu, d->n_conf_mu) - d->n_conf_mu);
}

typedef struct {
    const int32_t *x;          /* the gene's counts, N */
    const lpb200_design *d;
    const lpb200_model *m;
    const double *mat_drop;    /* for logistic_ofMu */
    int gene;
    const double *pi_fixed;    /* vecPiParam for "logistic" (fitMeanDispersion.R:887-892) or NULL */
    double *mu, *disp, *pi;    /* work, N each */
} genefit_ctx;

static double genefit_cost(int n, const double *th, void *ex) {
    (void)n;
    genefit_ctx *c = (genefit_ctx *)ex;
    const lpb200_design *d = c->d;
    int N = d->n_cells;
    gene_nat g;
    unpack_theta(d, c->m, th, &g);
    for (int j = 0; j < N; j++) {
        c->disp[j] = disp_at(d, c->m->disp_model, &g, j);
        c->mu[j] = mu_at(d, c->m->mu_model, &g, j);
    }
    const double *pi = NULL;
    if (c->pi_fixed) {
        pi = c->pi_fixed;
    } else if (c->m->drop_model != LPB200_DROP_NONE) { /* logistic_ofMu, :203-209 */
        for (int j = 0; j < N; j++) c->pi[j] = pi_at(d, c->m, c->mat_drop, c->gene, j, c->mu[j]);
        pi = c->pi;
    }
    return ll_gene(c->x, N, c->mu, d->size_factors, c->disp, pi);
}

/* fitMuDispGene (fitMeanDispersion.R:336-509): theta0 in optimiser coordinates */
static int fit_mudisp_gene(genefit_ctx *c, const double *theta0, const lpb200_control *ctl, gene_nat *out,
                           double *ll, int *conv, uint64_t *evals) {
    int n = n_theta(c->d, c->m);
    double b[LPB200_MAX_PARAMS];
    memcpy(b, theta0, sizeof(double) * n);
    optim_env e = {genefit_cost, c, ctl->ndeps, 0, 0};
    double fmin = NAN;
    int fc = 0, gc = 0, fail = 0;
    int rc = vmmin(n, b, &fmin, &e, NULL, ctl->maxit_mudisp, -INFINITY, ctl->reltol_mudisp, &fc, &gc, &fail);
    *evals += e.evals;
    if (rc != 0) {
        *conv = 1001;
        *ll = NAN;
        unpack_theta(c->d, c->m, theta0, out);
        return rc;
    }
    unpack_theta(c->d, c->m, b, out);
    *ll = fmin * (-1.0); /* value = Fmin * fnscale */
    *conv = fail;
    if (!isfinite(*ll)) *conv = 1001;
    return 0;
}

/* The fit objective itself, evalLogLikMuDispGeneFit (fitMeanDispersion.R:62-222), at caller-supplied optimiser
 * coordinates: theta is n_points x n_theta row-major.  For the tests that compare it with the reference's own R text
 * (tests/golden/reference_exec.json); the drop-out handling is fitMuDispGene's (:887-892 fixed pi for "logistic"). */
int lporacle_objective_mudisp(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                              const lpb200_params *p, int gene, const double *theta, int n_points, double *out) {
    if (!counts || !d || !m || !theta || !out || gene < 0 || gene >= d->n_genes) return LPB200_E_ARG;
    int N = d->n_cells, n = n_theta(d, m);
    double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)N);
    if (!buf) return LPB200_E_ARG;
    genefit_ctx gc;
    gc.x = counts + (size_t)gene * N;
    gc.d = d;
    gc.m = m;
    gc.mat_drop = p ? p->mat_drop : NULL;
    gc.gene = gene;
    gc.mu = buf;
    gc.disp = buf + N;
    gc.pi = buf + 2 * (size_t)N;
    gc.pi_fixed = NULL;
    if (m->drop_model == LPB200_DROP_LOGISTIC) {
        double *pf = buf + 3 * (size_t)N;
        for (int j = 0; j < N; j++) pf[j] = pi_at(d, m, p->mat_drop, gene, j, NAN);
        gc.pi_fixed = pf;
    }
    for (int k = 0; k < n_points; k++) out[k] = genefit_cost(n, theta + (size_t)k * n, &gc);
    free(buf);
    return 0;
}

/* pack natural-scale guesses into theta (fitMeanDispersion.R:349-364); for impulse the caller
 * supplies the 7 mu coordinates already in optimiser space (:621-622) */
static void pack_theta(const lpb200_design *d, const lpb200_model *m, const gene_nat *g,
                       const double *impulse_coords, double *th) {
    int u = 0;
    int pd = n_disp_par(m, d), pm = n_mu_par(m, d);
    for (int k = 0; k < pd; k++) th[u++] = (m->disp_model == LPB200_DISP_SPLINES) ? g->disp[k] : log(g->disp[k]);
    for (int c = 0; c < d->n_conf_disp; c++)
        for (int b = 1; b < d->n_batches_disp[c]; b++) th[u++] = log(g->bdisp[c][b]);
    if (m->mu_model == LPB200_MU_IMPULSE) {
        for (int k = 0; k < 7; k++) th[u++] = impulse_coords[k];
    } else {
        for (int k = 0; k < pm; k++) th[u++] = (m->mu_model == LPB200_MU_SPLINES) ? g->mu[k] : log(g->mu[k]);
    }
    for (int c = 0; c < d->n_conf_mu; c++)
        for (int b = 1; b < d->n_batches_mu[c]; b++) th[u++] = log(g->bmu[c][b]);
}

/* ------------------------------------------------------------------------- */
/* least squares y ~ 0 + B  (stats::lm): thin QR by twice-iterated Gram-Schmidt */
/* ------------------------------------------------------------------------- */

typedef struct {
    int n, k;
    double *Q; /* n x k column-major, orthonormal */
    double *R; /* k x k upper triangular, column-major */
} thinqr;

static int thinqr_build(const double *B, int n, int k, thinqr *qr) {
    qr->n = n;
    qr->k = k;
    qr->Q = (double *)malloc(sizeof(double) * (size_t)n * k);
    qr->R = (double *)calloc((size_t)k * k, sizeof(double));
    memcpy(qr->Q, B, sizeof(double) * (size_t)n * k);
    for (int c = 0; c < k; c++) {
        double *q = qr->Q + (size_t)c * n;
        for (int pass = 0; pass < 2; pass++) {
            for (int p = 0; p < c; p++) {
                const double *qp = qr->Q + (size_t)p * n;
                long double dot = 0.0L;
                for (int r = 0; r < n; r++) dot += (long double)qp[r] * q[r];
                for (int r = 0; r < n; r++) q[r] -= (double)dot * qp[r];
                qr->R[p + (size_t)c * k] += (double)dot;
            }
        }
        long double nrm = 0.0L;
        for (int r = 0; r < n; r++) nrm += (long double)q[r] * q[r];
        double nr = (double)sqrtl(nrm);
        if (!(nr > 0)) return -1;
        qr->R[c + (size_t)c * k] = nr;
        for (int r = 0; r < n; r++) q[r] /= nr;
    }
    return 0;
}
static void thinqr_free(thinqr *qr) {
    free(qr->Q);
    free(qr->R);
}
/* coef (k) and/or fitted (n) of y ~ 0 + B */
static void thinqr_solve(const thinqr *qr, const double *y, double *coef, double *fitted) {
    int n = qr->n, k = qr->k;
    double z[16];
    for (int c = 0; c < k; c++) {
        long double dot = 0.0L;
        const double *q = qr->Q + (size_t)c * n;
        for (int r = 0; r < n; r++) dot += (long double)q[r] * y[r];
        z[c] = (double)dot;
    }
    if (fitted) {
        for (int r = 0; r < n; r++) {
            double s = 0.0;
            for (int c = 0; c < k; c++) s += qr->Q[r + (size_t)c * n] * z[c];
            fitted[r] = s;
        }
    }
    if (coef) {
        for (int c = k - 1; c >= 0; c--) {
            double s = z[c];
            for (int p = c + 1; p < k; p++) s -= qr->R[c + (size_t)p * k] * coef[p];
            coef[c] = s / qr->R[c + (size_t)c * k];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* initialiseImpulseParameters (initialiseImpulseParameters.R:29-111)          */
/* ------------------------------------------------------------------------- */

static void impulse_starts_gene(const int32_t *x, const lpb200_design *d, const thinqr *qr, double *work /* 2N */,
                                double *peak, double *valley) {
    int N = d->n_cells;
    const double *t = d->pseudotime;
    double *y = work, *fit = work + N;
    for (int j = 0; j < N; j++) {
        double sf = d->size_factors ? d->size_factors[j] : 1.0;
        y[j] = log((double)(x[j] < 0 ? 0 : x[j]) / sf + 1); /* NA counts as 0 (processSCData.R:262-264) */
    }
    thinqr_solve(qr, y, NULL, fit);
    for (int j = 0; j < N; j++) fit[j] = exp(fit[j]);
    double tmin = t[0], tmax = t[0];
    for (int j = 1; j < N; j++) {
        if (t[j] < tmin) tmin = t[j];
        if (t[j] > tmax) tmax = t[j];
    }
    /* which.max / which.min return the FIRST extreme index */
    int imax_tail = 1, imin_tail = 1;     /* over fit[2:N] (0-based 1..N-1) */
    for (int j = 2; j < N; j++) {
        if (fit[j] > fit[imax_tail]) imax_tail = j;
        if (fit[j] < fit[imin_tail]) imin_tail = j;
    }
    int imax_head = 0, imin_head = 0;     /* over fit[1:(N-1)] */
    for (int j = 1; j < N - 1; j++) {
        if (fit[j] > fit[imax_head]) imax_head = j;
        if (fit[j] < fit[imin_head]) imin_head = j;
    }
    double fmax = fit[0], fmin = fit[0];
    for (int j = 1; j < N; j++) {
        if (fit[j] > fmax) fmax = fit[j];
        if (fit[j] < fmin) fmin = fit[j];
    }
    /* peak */
    double T1 = (t[0] + t[imax_tail]) / 2;
    double T2 = (t[imax_head] + t[N - 1]) / 2;
    if (T1 == t[0]) {
        double v = INFINITY;
        for (int j = 0; j < N; j++) if (t[j] != T1 && t[j] < v) v = t[j];
        T1 = v;
    }
    if (T2 == t[N - 1]) { /* :65-69 assigns scaT1Peak here -- reproduced */
        double v = -INFINITY;
        for (int j = 0; j < N; j++) if (t[j] != T2 && t[j] > v) v = t[j];
        T1 = v;
    }
    peak[0] = 2 * log(3.0) / (T1 - tmin);
    peak[1] = 2 * log(3.0) / (tmax - T2);
    peak[2] = log(fit[0] + 1);
    peak[3] = log(fmax + 1);
    peak[4] = log(fit[N - 1] + 1);
    peak[5] = T1;
    peak[6] = T2;
    /* valley */
    double V1 = (t[0] + t[imin_tail]) / 2;
    double V2 = (t[imin_head] + t[N - 1]) / 2;
    if (V1 == t[0]) {
        double v = INFINITY;
        for (int j = 0; j < N; j++) if (t[j] != V1 && t[j] < v) v = t[j];
        V1 = v;
    }
    if (V2 == t[N - 1]) {
        double v = -INFINITY;
        for (int j = 0; j < N; j++) if (t[j] != V2 && t[j] > v) v = t[j];
        V2 = v;
    }
    valley[0] = 2 * log(3.0) / (V1 - tmin);
    valley[1] = 2 * log(3.0) / (tmax - V2);
    valley[2] = log(fit[0] + 1);
    valley[3] = log(fmin + 1);
    valley[4] = log(fit[N - 1] + 1);
    valley[5] = V1;
    valley[6] = V2;
}

typedef struct {
    const int32_t *counts;
    const lpb200_design *d;
    const thinqr *qr;
    double *peak, *valley;
} starts_ctx;

static void starts_item(int i, void *vctx, void *scratch) {
    starts_ctx *c = (starts_ctx *)vctx;
    double pk[7], vl[7];
    int G = c->d->n_genes;
    impulse_starts_gene(c->counts + (size_t)i * c->d->n_cells, c->d, c->qr, (double *)scratch, pk, vl);
    for (int k = 0; k < 7; k++) {
        c->peak[i + (size_t)k * G] = pk[k];
        c->valley[i + (size_t)k * G] = vl[k];
    }
}

int lporacle_impulse_starts(const int32_t *counts, const lpb200_design *d, double *peak, double *valley,
                            int nthreads) {
    if (!d->impulse_basis || !d->pseudotime) return LPB200_E_ARG;
    thinqr qr;
    if (thinqr_build(d->impulse_basis, d->n_cells, 4, &qr)) return LPB200_E_ARG;
    starts_ctx c = {counts, d, &qr, peak, valley};
    run_pool(d->n_genes, nthreads, starts_item, &c, sizeof(double) * 2 * (size_t)d->n_cells);
    thinqr_free(&qr);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* fitMuDisp (fitMeanDispersion.R:853-1039) incl. fitMuDispGeneImpulse (:570-721) */
/* ------------------------------------------------------------------------- */

typedef struct {
    const int32_t *counts;
    const lpb200_design *d;
    const lpb200_model *m;
    lpb200_params *p;
    const lpb200_control *ctl;
    const thinqr *qr;
    double *ll;
    int32_t *conv;
    uint64_t evals; /* atomically added */
} fitmudisp_ctx;

static void fitmudisp_item(int i, void *vctx, void *scratch) {
    fitmudisp_ctx *c = (fitmudisp_ctx *)vctx;
    const lpb200_design *d = c->d;
    const lpb200_model *m = c->m;
    int N = d->n_cells;
    double *buf = (double *)scratch; /* mu, disp, pi, pi_fixed, work(2N) */
    genefit_ctx gc;
    gc.x = c->counts + (size_t)i * N;
    gc.d = d;
    gc.m = m;
    gc.mat_drop = c->p->mat_drop;
    gc.gene = i;
    gc.mu = buf;
    gc.disp = buf + N;
    gc.pi = buf + 2 * (size_t)N;
    gc.pi_fixed = NULL;
    if (m->drop_model == LPB200_DROP_LOGISTIC) { /* :887-892: constant during the gene's fit */
        double *pf = buf + 3 * (size_t)N;
        for (int j = 0; j < N; j++) pf[j] = pi_at(d, m, c->p->mat_drop, i, j, NAN);
        gc.pi_fixed = pf;
    }
    gene_nat guess, best;
    load_gene_nat(d, m, c->p, i, &guess);
    double th0[LPB200_MAX_PARAMS];
    uint64_t evals = 0;
    double ll = NAN;
    int conv = 0;
    if (m->mu_model == LPB200_MU_IMPULSE) {
        double pk[7], vl[7], prior[7];
        impulse_starts_gene(gc.x, d, c->qr, buf + 4 * (size_t)N, pk, vl);
        for (int k = 0; k < 7; k++) prior[k] = guess.mu[k];
        for (int k = 2; k < 5; k++) prior[k] = log(prior[k]); /* :621-622 */
        const double *starts[3] = {pk, vl, prior};            /* lsFits order Peak, Valley, Prior :663 */
        gene_nat fits[3];
        double lls[3];
        int convs[3];
        int any_nan = 0;
        for (int s = 0; s < 3; s++) {
            pack_theta(d, m, &guess, starts[s], th0);
            fit_mudisp_gene(&gc, th0, c->ctl, &fits[s], &lls[s], &convs[s], &evals);
            if (isnan(lls[s])) any_nan = 1;
        }
        if (!any_nan) { /* :664-671 first maximum */
            int bi = 0;
            for (int s = 1; s < 3; s++) if (lls[s] > lls[bi]) bi = s;
            best = fits[bi];
            ll = lls[bi];
            conv = convs[bi];
        } else { /* optim error => the reference aborts (fitMeanDispersion.R:384-389) */
            best = guess;
            ll = NAN;
            conv = 1001;
        }
    } else {
        pack_theta(d, m, &guess, NULL, th0);
        fit_mudisp_gene(&gc, th0, c->ctl, &best, &ll, &conv, &evals);
        if (conv == 1001) best = guess;
    }
    store_gene_nat(d, m, c->p, i, &best);
    c->ll[i] = ll;
    c->conv[i] = conv;
    __atomic_fetch_add(&c->evals, evals, __ATOMIC_RELAXED);
}

int lporacle_fit_mudisp(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                        lpb200_params *p, const lpb200_control *ctl, double *ll, int32_t *conv,
                        uint64_t *fn_evals, int nthreads) {
    if (m->mu_model == LPB200_MU_MM) return LPB200_E_NOTSUP; /* fitMeanDispersion.R:804 stop() */
    if (n_mu_par(m, d) < 0 || n_disp_par(m, d) < 0 || n_theta(d, m) > LPB200_MAX_PARAMS) return LPB200_E_ARG;
    thinqr qr;
    int have_qr = 0;
    if (m->mu_model == LPB200_MU_IMPULSE) {
        if (!d->impulse_basis || !d->pseudotime) return LPB200_E_ARG;
        if (thinqr_build(d->impulse_basis, d->n_cells, 4, &qr)) return LPB200_E_ARG;
        have_qr = 1;
    }
    fitmudisp_ctx c = {counts, d, m, p, ctl, have_qr ? &qr : NULL, ll, conv, 0};
    run_pool(d->n_genes, nthreads, fitmudisp_item, &c, sizeof(double) * 6 * (size_t)d->n_cells);
    if (have_qr) thinqr_free(&qr);
    if (fn_evals) *fn_evals += c.evals;
    for (int i = 0; i < d->n_genes; i++)
        if (conv[i] == 1001) return LPB200_E_NONFINITE;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* fitPi (fitDropout.R:418-515)                                               */
/* ------------------------------------------------------------------------- */

/* theta in optimiser coordinates -> linear model (fitDropout.R:72-81 == :302-311) */
static void pi_theta_transform(int drop_model, int q, const double *th, double *lin) {
    const double HI = pow(10.0, 10.0), TINY = pow(10.0, -10.0);
    for (int k = 0; k < q; k++) lin[k] = th[k];
    if (drop_model == LPB200_DROP_LOGISTIC_OFMU) lin[1] = -exp(lin[1]);
    for (int k = 0; k < q; k++) {
        if (lin[k] < -HI) lin[k] = -HI;
        if (lin[k] > HI) lin[k] = HI;
    }
    if (drop_model == LPB200_DROP_LOGISTIC_OFMU)
        if (lin[1] > -TINY) lin[1] = -TINY;
}

typedef struct {
    int G, q, drop_model;
    const int32_t *counts; /* gene-major; cell j's column is strided */
    int N, j;
    double sf;
    const double *mu, *disp;   /* G each: mu_ij (no size factor), phi_ij */
    const double *pred;        /* G x q row-major: matPiAllPredictors */
} cellfit_ctx;

/* evalLogLikPiZINB_SingleCell (fitDropout.R:59-100) */
static double cellfit_cost(int n, const double *th, void *ex) {
    (void)n;
    cellfit_ctx *c = (cellfit_ctx *)ex;
    double lin[2 + LPB200_MAX_PI_PRED];
    pi_theta_transform(c->drop_model, c->q, th, lin);
    long double sz = 0.0L, snz = 0.0L;
    for (int i = 0; i < c->G; i++) {
        int32_t x = c->counts[(size_t)i * c->N + c->j];
        if (x < 0) continue;
        double pi = lporacle_dropout_rate(lin, c->pred + (size_t)i * c->q, c->q);
        double v = lporacle_ll_zinb_obs((double)x, c->mu[i] * c->sf, c->disp[i], pi);
        if (x == 0) sz += v; else snz += v;
    }
    return (double)sz + (double)snz;
}

typedef struct {
    const int32_t *counts;
    const lpb200_design *d;
    const lpb200_model *m;
    lpb200_params *p;
    const lpb200_control *ctl;
    const gene_nat *gn;  /* G gene models */
    double *new_drop;    /* N x q */
    double *ll;
    int32_t *conv;
    uint64_t evals;
} fitpi_ctx;

static void fitpi_item(int j, void *vctx, void *scratch) {
    fitpi_ctx *c = (fitpi_ctx *)vctx;
    const lpb200_design *d = c->d;
    const lpb200_model *m = c->m;
    int G = d->n_genes, N = d->n_cells;
    int q = n_drop_par(m, d);
    double *mu = (double *)scratch, *disp = mu + G, *pred = disp + G;
    for (int i = 0; i < G; i++) { /* fitDropout.R:432-455 */
        mu[i] = mu_at(d, m->mu_model, &c->gn[i], j);
        disp[i] = disp_at(d, m->disp_model, &c->gn[i], j);
        int u = 0;
        pred[(size_t)i * q + u++] = 1.0;
        if (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) pred[(size_t)i * q + u++] = log(mu[i]);
        for (int k = 0; k < d->n_pi_pred; k++) pred[(size_t)i * q + u++] = d->pi_const_pred[i + (size_t)k * G];
    }
    double th[2 + LPB200_MAX_PI_PRED];
    for (int k = 0; k < q; k++) th[k] = c->p->mat_drop[j + (size_t)k * N];
    if (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) th[1] = log(-th[1]); /* :463-466 */
    cellfit_ctx cc = {G, q, m->drop_model, c->counts, N, j, d->size_factors ? d->size_factors[j] : 1.0, mu, disp, pred};
    optim_env e = {cellfit_cost, &cc, c->ctl->ndeps, 0, 0};
    double fmin = NAN;
    int fc = 0, gc = 0, fail = 0;
    int rc = vmmin(q, th, &fmin, &e, NULL, c->ctl->maxit_pi, -INFINITY, c->ctl->reltol_pi, &fc, &gc, &fail);
    __atomic_fetch_add(&c->evals, e.evals, __ATOMIC_RELAXED);
    if (rc != 0) {
        c->conv[j] = 1001;
        c->ll[j] = NAN;
        for (int k = 0; k < q; k++) c->new_drop[j + (size_t)k * N] = c->p->mat_drop[j + (size_t)k * N];
        return;
    }
    double lin[2 + LPB200_MAX_PI_PRED];
    pi_theta_transform(m->drop_model, q, th, lin);
    for (int k = 0; k < q; k++) c->new_drop[j + (size_t)k * N] = lin[k];
    c->conv[j] = fail;
    c->ll[j] = fmin * (-1.0);
}

/* evalLogLikPiZINB_ManyCells (fitDropout.R:169-207): one theta for all cells */
typedef struct {
    const int32_t *counts;
    const lpb200_design *d;
    const lpb200_model *m;
    lpb200_params *p;  /* mat_drop is overwritten with the tiled theta */
    int nthreads;
    double *llbuf;
} manycells_ctx;

static double manycells_cost(int n, const double *th, void *ex) {
    (void)n;
    manycells_ctx *c = (manycells_ctx *)ex;
    int q = n_drop_par(c->m, c->d), N = c->d->n_cells, G = c->d->n_genes;
    double lin[2 + LPB200_MAX_PI_PRED];
    pi_theta_transform(c->m->drop_model, q, th, lin);
    for (int k = 0; k < q; k++)
        for (int j = 0; j < N; j++) c->p->mat_drop[j + (size_t)k * N] = lin[k];
    lporacle_eval_loglik(c->counts, c->d, c->m, c->p, c->llbuf, c->nthreads);
    long double s = 0.0L;
    for (int i = 0; i < G; i++) s += c->llbuf[i];
    return (double)s;
}

int lporacle_fit_pi(const int32_t *counts, const lpb200_design *d, const lpb200_model *m, lpb200_params *p,
                    const lpb200_control *ctl, double *ll, int32_t *conv, uint64_t *fn_evals, int nthreads) {
    if (m->drop_model == LPB200_DROP_NONE) return LPB200_E_ARG;
    if (m->mu_model == LPB200_MU_MM) return LPB200_E_NOTSUP; /* fitDropout.R:429-431 */
    int G = d->n_genes, N = d->n_cells, q = n_drop_par(m, d);
    if (m->drop_fit_group == LPB200_DROPFIT_PERCELL) {
        gene_nat *gn = (gene_nat *)malloc(sizeof(gene_nat) * (size_t)G);
        for (int i = 0; i < G; i++) load_gene_nat(d, m, p, i, &gn[i]);
        double *nd = (double *)malloc(sizeof(double) * (size_t)N * q);
        fitpi_ctx c = {counts, d, m, p, ctl, gn, nd, ll, conv, 0};
        run_pool(N, nthreads, fitpi_item, &c, sizeof(double) * (size_t)G * (2 + q));
        memcpy(p->mat_drop, nd, sizeof(double) * (size_t)N * q);
        free(nd);
        free(gn);
        if (fn_evals) *fn_evals += c.evals;
        for (int j = 0; j < N; j++)
            if (conv[j] == 1001) return LPB200_E_NONFINITE;
        return 0;
    }
    /* ForAllCells, fitDropout.R:480-499 */
    double th[2 + LPB200_MAX_PI_PRED];
    for (int k = 0; k < q; k++) th[k] = p->mat_drop[0 + (size_t)k * N]; /* all rows are the same */
    if (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) th[1] = log(-th[1]);
    double *llbuf = (double *)malloc(sizeof(double) * (size_t)G);
    manycells_ctx mc = {counts, d, m, p, nthreads, llbuf};
    optim_env e = {manycells_cost, &mc, ctl->ndeps, 0, 0};
    double fmin = NAN;
    int fc = 0, gc = 0, fail = 0;
    int rc = vmmin(q, th, &fmin, &e, NULL, ctl->maxit_pi, -INFINITY, ctl->reltol_pi, &fc, &gc, &fail);
    free(llbuf);
    if (fn_evals) *fn_evals += e.evals * (uint64_t)G;
    if (rc != 0) {
        conv[0] = 1001;
        ll[0] = NAN;
        return rc;
    }
    double lin[2 + LPB200_MAX_PI_PRED];
    pi_theta_transform(m->drop_model, q, th, lin);
    for (int k = 0; k < q; k++)
        for (int j = 0; j < N; j++) p->mat_drop[j + (size_t)k * N] = lin[k];
    conv[0] = fail;
    ll[0] = fmin * (-1.0);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* initial values (initialiseMuModel.R:93-133,160-179; initialiseDispModel.R:89-130;   */
/* initialiseDropModel.R:73-91)                                                */
/* ------------------------------------------------------------------------- */

int lporacle_row_means(const int32_t *counts, const lpb200_design *d, double *means) {
    int G = d->n_genes, N = d->n_cells;
    for (int i = 0; i < G; i++) { /* Matrix::rowMeans(na.rm = TRUE) */
        double s = 0.0;
        int n = 0;
        for (int j = 0; j < N; j++) {
            int32_t x = counts[(size_t)i * N + j];
            if (x >= 0) { s += x; n++; }
        }
        means[i] = s / n;
    }
    return 0;
}

int lporacle_init_params(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         lpb200_params *out, double min_rowmean_global) {
    int G = d->n_genes, N = d->n_cells;
    if (m->mu_model == LPB200_MU_MM) return LPB200_E_NOTSUP;
    int pm = n_mu_par(m, d), pd = n_disp_par(m, d);
    if (pm < 0 || pd < 0) return LPB200_E_ARG;
    double *means = (double *)malloc(sizeof(double) * (size_t)G);
    lporacle_row_means(counts, d, means);
    double minmean = INFINITY;
    for (int i = 0; i < G; i++) if (means[i] < minmean) minmean = means[i];
    if (!isnan(min_rowmean_global)) minmean = min_rowmean_global;
    for (int i = 0; i < G; i++) {
        double mi = means[i];
        if (mi < 1e-10) mi = 1e-10; /* initialiseMuModel.R:95 */
        if (m->mu_model == LPB200_MU_CONSTANT) {
            out->mat_mu[i] = mi;
        } else if (m->mu_model == LPB200_MU_IMPULSE) { /* (1,1,m,m,m,1,1) :99-102 */
            for (int k = 0; k < 7; k++) out->mat_mu[i + (size_t)k * G] = 1.0;
            for (int k = 2; k < 5; k++) out->mat_mu[i + (size_t)k * G] = mi;
        } else if (m->mu_model == LPB200_MU_GROUPS) {
            for (int k = 0; k < pm; k++) out->mat_mu[i + (size_t)k * G] = mi;
        }
    }
    if (m->mu_model == LPB200_MU_SPLINES) { /* lm(log(x+1) ~ 0 + ns(...))$coef :111-118 */
        thinqr qr;
        if (thinqr_build(d->spline_basis_mu, N, pm, &qr)) { free(means); return LPB200_E_ARG; }
        double *y = (double *)malloc(sizeof(double) * (size_t)N);
        double coef[16];
        for (int i = 0; i < G; i++) {
            for (int j = 0; j < N; j++) {
                int32_t xv = counts[(size_t)i * N + j];
                y[j] = log((double)(xv < 0 ? 0 : xv) + 1); /* NA counts as 0 */
            }
            thinqr_solve(&qr, y, coef, NULL);
            for (int k = 0; k < pm; k++) out->mat_mu[i + (size_t)k * G] = coef[k];
        }
        free(y);
        thinqr_free(&qr);
    }
    for (int i = 0; i < G; i++) { /* initialiseDispModel.R:90-103 */
        for (int k = 0; k < pd; k++)
            out->mat_disp[i + (size_t)k * G] = (m->disp_model == LPB200_DISP_SPLINES && k > 0) ? 0.0 : 1.0;
    }
    if (d->n_conf_mu > 0) {
        size_t tot = (size_t)G * n_batch_cols(d->n_batches_mu, d->n_conf_mu);
        for (size_t k = 0; k < tot; k++) out->batch_mu[k] = 1.0;
    }
    if (d->n_conf_disp > 0) {
        size_t tot = (size_t)G * n_batch_cols(d->n_batches_disp, d->n_conf_disp);
        for (size_t k = 0; k < tot; k++) out->batch_disp[k] = 1.0;
    }
    if (m->drop_model != LPB200_DROP_NONE && out->mat_drop) {
        int q = n_drop_par(m, d);
        double target = 0.99;
        for (int k = 0; k < q; k++)
            for (int j = 0; j < N; j++) out->mat_drop[j + (size_t)k * N] = 0.0;
        if (m->drop_model == LPB200_DROP_LOGISTIC_OFMU) {
            double slope = -1;
            double off = log(target) - log(1 - target) - slope * log(minmean);
            for (int j = 0; j < N; j++) {
                out->mat_drop[j] = off;
                out->mat_drop[j + (size_t)N] = slope;
            }
        } else {
            double off = log(target) - log(1 - target);
            for (int j = 0; j < N; j++) out->mat_drop[j] = off;
        }
    }
    free(means);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* fitModel (fitModel.R:137-356)                                              */
/* ------------------------------------------------------------------------- */

int lporacle_fit_model(const int32_t *counts, const lpb200_design *d, const lpb200_model *m, lpb200_params *p,
                       int fit_drop, const lpb200_control *ctl, double min_rowmean_global, double *em_ll,
                       int32_t *n_iter, int32_t *converged, double *ll_init, double *ll, int32_t *conv,
                       uint64_t *fn_evals, int nthreads) {
    int G = d->n_genes, N = d->n_cells;
    int max_cycles = ctl->max_cycles;
    if (!fit_drop) max_cycles = 1; /* :181-183 */
    for (int k = 0; k < (fit_drop ? ctl->max_cycles : 1); k++) em_ll[k] = NAN;
    lpb200_params init = *p;
    if (!fit_drop) init.mat_drop = NULL; /* supplied lsDropModel is kept (initialiseDropModel.R:60) */
    int rc = lporacle_init_params(counts, d, m, &init, min_rowmean_global);
    if (rc) return rc;
    rc = lporacle_eval_loglik(counts, d, m, p, ll, nthreads); /* :228-233 */
    if (rc) return rc;
    long double s = 0.0L;
    for (int i = 0; i < G; i++) s += ll[i];
    double ll_new = (double)s, ll_old = NAN;
    if (ll_init) *ll_init = ll_new;
    int iter = 1;
    double *pill = (double *)malloc(sizeof(double) * (size_t)N);
    int32_t *piconv = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    while (iter == 1 || (ll_new > ll_old * ctl->prec_em && iter <= max_cycles)) { /* :242-243 */
        if (fit_drop) {
            rc = lporacle_fit_pi(counts, d, m, p, ctl, pill, piconv, fn_evals, nthreads);
            if (rc) break;
        }
        rc = lporacle_fit_mudisp(counts, d, m, p, ctl, ll, conv, fn_evals, nthreads);
        if (rc) break;
        ll_old = ll_new;
        s = 0.0L;
        for (int i = 0; i < G; i++) s += ll[i];
        ll_new = (double)s;
        em_ll[iter - 1] = ll_new;
        iter++;
    }
    free(pill);
    free(piconv);
    if (rc) return rc;
    int all0 = 1;
    for (int i = 0; i < G; i++) if (conv[i] != 0) all0 = 0;
    *converged = (all0 && ll_new < ll_old * ctl->prec_em && ll_new > ll_old) ? 1 : 0; /* :342-347 */
    *n_iter = iter - 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* calcPostDrop_Matrix (calcPostDrop.R:69-125); mu WITHOUT size factor (:110)  */
/* ------------------------------------------------------------------------- */

int lporacle_post_drop(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                       const lpb200_params *p, const int32_t *gene_ids, int n_ids, double *z) {
    int N = d->n_cells;
    if (m->drop_model == LPB200_DROP_NONE) return LPB200_E_ARG;
    for (int r = 0; r < n_ids; r++) {
        int i = gene_ids[r];
        gene_nat g;
        load_gene_nat(d, m, p, i, &g);
        for (int j = 0; j < N; j++) {
            double mu = mu_at(d, m->mu_model, &g, j);
            double phi = disp_at(d, m->disp_model, &g, j);
            double pi = pi_at(d, m, p->mat_drop, i, j, mu);
            double nb0 = pow(phi / (phi + mu), phi);
            double zz = pi / (pi + (1 - pi) * nb0);
            int32_t x = counts[(size_t)i * N + j];
            z[r + (size_t)j * n_ids] = (x < 0) ? NAN : (x > 0 ? 0.0 : zz);
        }
    }
    return 0;
}

/* getFits.R:52-363: getFitsMean / getFitsDispersion / getFitsDropout / getNormData */
int lporacle_expand_fits(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         const lpb200_params *p, const int32_t *gene_ids, int n_ids, int what, double *z) {
    int N = d->n_cells;
    int kind = what & 15;
    if (kind == LPB200_EXPAND_POSTDROP) return lporacle_post_drop(counts, d, m, p, gene_ids, n_ids, z);
    for (int r = 0; r < n_ids; r++) {
        int i = gene_ids[r];
        gene_nat g;
        load_gene_nat(d, m, p, i, &g);
        for (int j = 0; j < N; j++) {
            double out;
            if (kind == LPB200_EXPAND_MEAN) out = mu_at(d, m->mu_model, &g, j);
            else if (kind == LPB200_EXPAND_DISP) out = disp_at(d, m->disp_model, &g, j);
            else if (kind == LPB200_EXPAND_DROP) out = pi_at(d, m, p->mat_drop, i, j, mu_at(d, m->mu_model, &g, j));
            else { /* getNormData */
                int32_t x = counts[(size_t)i * N + j];
                if (x < 0) out = NAN;
                else {
                    out = (double)x;
                    if (!(what & LPB200_EXPAND_NO_DEPTH)) out = out / (d->size_factors ? d->size_factors[j] : 1.0);
                    if (!(what & LPB200_EXPAND_NO_BATCH)) {
                        double bp = 1.0;
                        for (int c = 0; c < d->n_conf_mu; c++) bp = bp * g.bmu[c][d->batch_idx_mu[(size_t)c * N + j]];
                        out = out / bp;
                    }
                }
            }
            z[r + (size_t)j * n_ids] = out;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* LRT + BH (runHypothesisTests.R:81-86)                                      */
/* ------------------------------------------------------------------------- */

typedef struct { double p; int idx; } pidx;
static int cmp_desc(const void *a, const void *b) {
    const pidx *x = (const pidx *)a, *y = (const pidx *)b;
    if (x->p > y->p) return -1;
    if (x->p < y->p) return 1;
    return x->idx - y->idx;
}

int lporacle_lrt(const double *ll_full, const double *ll_red, int ddf, int G, double *p, double *padj) {
    pidx *o = (pidx *)malloc(sizeof(pidx) * (size_t)(G > 0 ? G : 1));
    int n = 0;
    for (int i = 0; i < G; i++) {
        double dev = 2 * (ll_full[i] - ll_red[i]);
        p[i] = lporacle_pchisq_upper(dev, (double)ddf);
        padj[i] = NAN;
        if (!isnan(p[i])) { o[n].p = p[i]; o[n].idx = i; n++; }
    }
    qsort(o, n, sizeof(pidx), cmp_desc);
    /* p.adjust BH: pmin(1, cummin(n/i * p[o]))[ro], i = n..1 over decreasing p */
    double cm = INFINITY;
    for (int k = 0; k < n; k++) {
        double v = (double)n / (double)(n - k) * o[k].p;
        if (v < cm) cm = v;
        padj[o[k].idx] = cm < 1.0 ? cm : 1.0;
    }
    free(o);
    return 0;
}
