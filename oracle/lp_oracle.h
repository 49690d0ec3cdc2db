/*
 * lp_oracle.h -- CPU oracle: a plain-C restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in lineagepulse_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / reported CPU baseline.
 *
 * PARITY STATUS: the reference ships no tests, fixtures or golden vectors
 *  and R cannot be run here (F3): "parity unpinned" by an R session.
 * The pins this oracle is checked against: the reference's own R source text
 * executed by tests/r_eval.py (outputs committed as tests/golden/reference_exec.json
 * and reproduced to the last bit), and for the R-base builtins that evaluator
 * restates: 50-digit mpmath and scipy values (tests/golden/), the ?optim Rosenbrock
 * known answers, and analytic identities.
 */
#ifndef LP_ORACLE_H
#define LP_ORACLE_H
#include "../include/lpb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* counts: gene-major G x N int32 (row i contiguous), LPB200_NA_COUNT = NA.  G, N from the design. */

int lporacle_row_means(const int32_t *counts, const lpb200_design *d, double *means);
int lporacle_init_params(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         lpb200_params *out, double min_rowmean_global);
int lporacle_eval_loglik(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         const lpb200_params *p, double *ll, int nthreads);
int lporacle_impulse_starts(const int32_t *counts, const lpb200_design *d, double *peak,
                            double *valley, int nthreads);
int lporacle_fit_mudisp(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                        lpb200_params *p, const lpb200_control *c, double *ll, int32_t *conv,
                        uint64_t *fn_evals, int nthreads);
/* evalLogLikMuDispGeneFit (fitMeanDispersion.R:62-222) of one gene at n_points parameter vectors (row-major) */
int lporacle_objective_mudisp(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                              const lpb200_params *p, int gene, const double *theta, int n_points, double *out);
int lporacle_fit_pi(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                    lpb200_params *p, const lpb200_control *c, double *ll, int32_t *conv,
                    uint64_t *fn_evals, int nthreads);
int lporacle_fit_model(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                       lpb200_params *p, int fit_drop, const lpb200_control *c,
                       double min_rowmean_global, double *em_ll, int32_t *n_iter,
                       int32_t *converged, double *ll_init, double *ll, int32_t *conv,
                       uint64_t *fn_evals, int nthreads);
int lporacle_post_drop(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                       const lpb200_params *p, const int32_t *gene_ids, int n_ids, double *z);
int lporacle_expand_fits(const int32_t *counts, const lpb200_design *d, const lpb200_model *m,
                         const lpb200_params *p, const int32_t *gene_ids, int n_ids, int what, double *z);
int lporacle_lrt(const double *ll_full, const double *ll_red, int ddf, int G, double *p, double *padj);

/* 0 (default): long double sums as R on x86; 1: double sums (R without long double) */
void lporacle_set_sum_mode(int mode);

/* scalar pieces exposed for known-answer tests */
double lporacle_dnbinom_mu_log(double x, double size, double mu);
double lporacle_pchisq_upper(double q, double df);
double lporacle_impulse_value(const double *par7, double t);
double lporacle_dropout_rate(const double *theta, const double *pred, int q);
double lporacle_ll_zinb_obs(double x, double mu, double disp, double pi); /* pi < 0 => NB */
/* optim(method="BFGS") on the ?optim Rosenbrock example; use_numgrad: gr=NULL path */
int lporacle_vmmin_rosenbrock(int use_numgrad, int maxit, double reltol, double *par, double *value,
                              int *fncount, int *grcount, int *conv);

#ifdef __cplusplus
}
#endif
#endif
