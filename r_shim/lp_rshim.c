/*
 * lp_rshim.c -- the `.Call` shim between the LineagePulse R package and liblpb200.so.
 *
 * R stays the host language (BASELINE.json north_star); this file contains no arithmetic: it
 * unpacks the R lists that fitMuDisp / fitPi / evalLogLikMatrix / fitModel already receive
 * (fitMeanDispersion.R:853-859, fitDropout.R:418-422, evalLogLik.R:307-313, fitModel.R:137-158),
 * calls the C ABI of include/lpb200.h and re-packs the results under the reference's names.
 *
 * STATUS: R is not installed in the build image.  This file is compiled against r_shim/fake_R/Rinternals.h
 * (a stand-in declaring only the R API used here) and, linked with the functional stand-in runtime
 * tests/csrc/fake_r_runtime.c, it is RUN on the GPU by tests/test_gpu_parity.py::test_r_shim_runs_on_the_gpu
 * (lp_create -> lp_load_counts_csc -> lp_set_design -> lp_fit_model -> lp_fit_models -> lp_eval_loglik -> lp_lrt,
 * compared bit for bit with the ctypes path); it has NOT been run against a real R.
 * Build inside the package with:  PKG_LIBS = -L<dir> -llpb200   (src/Makevars)
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "../include/lpb200.h"

/* ---- helpers -------------------------------------------------------------------------------- */
static SEXP list_get(SEXP list, const char *name) {
    SEXP names = getAttrib(list, R_NamesSymbol);
    for (int i = 0; i < LENGTH(list); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}
static const double *dbl_or_null(SEXP x) { return (x == R_NilValue) ? NULL : REAL(x); }

static void ctx_finalizer(SEXP ptr) {
    lpb200_ctx *ctx = (lpb200_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx) {
        lpb200_destroy(ctx);
        R_ClearExternalPtr(ptr);
    }
}
static lpb200_ctx *ctx_of(SEXP ptr) {
    lpb200_ctx *ctx = (lpb200_ctx *)R_ExternalPtrAddr(ptr);
    if (!ctx) error("LineagePulse: GPU context was released");
    return ctx;
}
static void check(lpb200_ctx *ctx, int rc) { /* any optimiser error aborts, as fitMeanDispersion.R:384-389 */
    if (rc != 0) error("LineagePulse (lpb200, code %d): %s", rc, lpb200_last_error(ctx));
}

/* ---- optimiser constants --------------------------------------------------------------------
 * The reference keeps them in the model lists (lsMuModelGlobal$MAXIT_BFGS_MuDisp / RELTOL_BFGS_MuDisp,
 * fitMeanDispersion.R:912-913; lsDropModelGlobal$MAXIT_BFGS_Pi / RELTOL_BFGS_Pi, fitDropout.R:291-292) and
 * fitModel's scaMaxEstimationCycles (fitModel.R:156,243).  The R wrappers send them before a fit with
 * .Call("lp_set_control", c(maxit_mudisp, reltol_mudisp, maxit_pi, reltol_pi, max_cycles)); NA entries and a
 * shorter vector leave fitModel.R:164-175's constants.  R calls .Call from one thread: a static is enough. */
static lpb200_control g_control;
static int g_control_set = 0;
static const lpb200_control *current_control(void) {
    if (!g_control_set) {
        lpb200_default_control(&g_control);
        g_control_set = 1;
    }
    return &g_control;
}
SEXP lp_set_control(SEXP v) {
    lpb200_default_control(&g_control);
    g_control_set = 1;
    if (v != R_NilValue) {
        int n = LENGTH(v);
        const double *x = REAL(v);
        if (n > 0 && !ISNAN(x[0])) g_control.maxit_mudisp = (int)x[0];
        if (n > 1 && !ISNAN(x[1])) g_control.reltol_mudisp = x[1];
        if (n > 2 && !ISNAN(x[2])) g_control.maxit_pi = (int)x[2];
        if (n > 3 && !ISNAN(x[3])) g_control.reltol_pi = x[3];
        if (n > 4 && !ISNAN(x[4])) g_control.max_cycles = (int)x[4];
        if (g_control.maxit_mudisp < 0 || g_control.maxit_pi < 0 || g_control.max_cycles < 1 ||
            !(g_control.reltol_mudisp >= 0) || !(g_control.reltol_pi >= 0)) {
            lpb200_default_control(&g_control);
            error("LineagePulse: lp_set_control: bad optimiser constants");
        }
    }
    return R_NilValue;
}

/* ---- context + data ------------------------------------------------------------------------- */
/* .Call("lp_create", devices) -> external pointer with finalizer.  devices: integer vector of CUDA device ids;
 * one id = one GPU, several = one gene-sharded context over all of them (lpb200_create_multi): what scaNProc /
 * MulticoreParam(workers = scaNProc) was for the CPU path (main_LineagePulse.R:271-275). */
SEXP lp_create(SEXP devices) {
    int n = LENGTH(devices);
    if (n < 1) error("LineagePulse: lp_create needs at least one device id");
    lpb200_ctx *ctx = (n == 1) ? lpb200_create(INTEGER(devices)[0]) : lpb200_create_multi(INTEGER(devices), n);
    if (!ctx) error("LineagePulse: no usable CUDA device (there is no CPU fallback)");
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call("lp_load_counts_csc", ctx, dim, p, i, x): slots of the dgCMatrix matCountsProc */
SEXP lp_load_counts_csc(SEXP sctx, SEXP dim, SEXP p, SEXP i, SEXP x) {
    lpb200_ctx *ctx = ctx_of(sctx);
    check(ctx, lpb200_load_counts_csc(ctx, INTEGER(dim)[0], INTEGER(dim)[1], INTEGER(p), INTEGER(i), REAL(x)));
    return R_NilValue;
}
/* .Call("lp_load_counts_dense", ctx, matrix) */
SEXP lp_load_counts_dense(SEXP sctx, SEXP mat) {
    lpb200_ctx *ctx = ctx_of(sctx);
    SEXP dim = getAttrib(mat, R_DimSymbol);
    check(ctx, lpb200_load_counts_dense(ctx, INTEGER(dim)[0], INTEGER(dim)[1], REAL(mat)));
    return R_NilValue;
}

/* the design travels as a named list built by the R wrapper from lsMuModelGlobal / lsDispModelGlobal /
 * lsDropModel (R/lpb200_wrappers.R::lpDesign); index vectors arrive 1-based and are shifted here */
static int *shift_idx(SEXP v) {
    if (v == R_NilValue) return NULL;
    int n = LENGTH(v);
    int *out = (int *)R_alloc(n, sizeof(int));
    for (int k = 0; k < n; k++) out[k] = INTEGER(v)[k] - 1;
    return out;
}
static void fill_design(SEXP lst, lpb200_design *d) {
    memset(d, 0, sizeof(*d));
    d->n_genes = asInteger(list_get(lst, "n_genes"));
    d->n_cells = asInteger(list_get(lst, "n_cells"));
    d->size_factors = dbl_or_null(list_get(lst, "vecNormConst"));
    d->pseudotime = dbl_or_null(list_get(lst, "vecContinuousCovar"));
    d->group_idx = shift_idx(list_get(lst, "vecidxGroups"));
    d->n_groups = asInteger(list_get(lst, "scaNGroups"));
    SEXP b = list_get(lst, "matSplineBasisMu");
    if (b != R_NilValue) { d->spline_basis_mu = REAL(b); d->k_spline_mu = INTEGER(getAttrib(b, R_DimSymbol))[1]; }
    b = list_get(lst, "matSplineBasisDisp");
    if (b != R_NilValue) { d->spline_basis_disp = REAL(b); d->k_spline_disp = INTEGER(getAttrib(b, R_DimSymbol))[1]; }
    d->impulse_basis = dbl_or_null(list_get(lst, "matImpulseBasis")); /* ns(t, df = 4, intercept = TRUE) */
    SEXP nb = list_get(lst, "vecNBatchesMu");
    if (nb != R_NilValue) {
        d->n_conf_mu = LENGTH(nb);
        for (int c = 0; c < d->n_conf_mu; c++) d->n_batches_mu[c] = INTEGER(nb)[c];
        d->batch_idx_mu = shift_idx(list_get(lst, "vecidxBatchAssignMu")); /* unlist(lsvecidxBatchAssign) */
    }
    nb = list_get(lst, "vecNBatchesDisp");
    if (nb != R_NilValue) {
        d->n_conf_disp = LENGTH(nb);
        for (int c = 0; c < d->n_conf_disp; c++) d->n_batches_disp[c] = INTEGER(nb)[c];
        d->batch_idx_disp = shift_idx(list_get(lst, "vecidxBatchAssignDisp"));
    }
    b = list_get(lst, "matPiConstPredictors");
    if (b != R_NilValue) { d->pi_const_pred = REAL(b); d->n_pi_pred = INTEGER(getAttrib(b, R_DimSymbol))[1]; }
}
SEXP lp_set_design(SEXP sctx, SEXP lst) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_design d;
    fill_design(lst, &d);
    check(ctx, lpb200_set_design(ctx, &d));
    return R_NilValue;
}

static lpb200_model model_of(SEXP m) { /* integer(4): mu, disp, drop, drop-fit-group codes */
    lpb200_model mm = {INTEGER(m)[0], INTEGER(m)[1], INTEGER(m)[2], INTEGER(m)[3]};
    return mm;
}
/* params: list(matMuModel, matBatchMu (cbind of lsmatBatchModel), matDispModel, matBatchDisp, matDropoutLinModel);
 * the matrices are duplicated by the R wrapper before the call, so writing into them is safe */
static lpb200_params params_of(SEXP p) {
    lpb200_params pp;
    pp.mat_mu = (double *)dbl_or_null(list_get(p, "matMuModel"));
    pp.batch_mu = (double *)dbl_or_null(list_get(p, "matBatchMu"));
    pp.mat_disp = (double *)dbl_or_null(list_get(p, "matDispModel"));
    pp.batch_disp = (double *)dbl_or_null(list_get(p, "matBatchDisp"));
    pp.mat_drop = (double *)dbl_or_null(list_get(p, "matDropoutLinModel"));
    return pp;
}

/* ---- evalLogLikMatrix (evalLogLik.R:307) ---------------------------------------------------- */
SEXP lp_eval_loglik(SEXP sctx, SEXP model, SEXP params, SEXP n_genes) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_model m = model_of(model);
    lpb200_params p = params_of(params);
    SEXP ll = PROTECT(allocVector(REALSXP, asInteger(n_genes)));
    check(ctx, lpb200_eval_loglik(ctx, &m, &p, REAL(ll)));
    UNPROTECT(1);
    return ll;
}

/* ---- fitMuDisp (fitMeanDispersion.R:853): returns list(params, vecConvergence, vecLL) -------- */
SEXP lp_fit_mudisp(SEXP sctx, SEXP model, SEXP params, SEXP n_genes) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_model m = model_of(model);
    lpb200_params p = params_of(params);
    lpb200_control c = *current_control();
    int G = asInteger(n_genes);
    SEXP ll = PROTECT(allocVector(REALSXP, G)), conv = PROTECT(allocVector(INTSXP, G));
    check(ctx, lpb200_fit_mudisp(ctx, &m, &p, &c, REAL(ll), INTEGER(conv)));
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, params);
    SET_VECTOR_ELT(out, 1, conv);
    SET_VECTOR_ELT(out, 2, ll);
    UNPROTECT(3);
    return out;
}

/* ---- fitPi (fitDropout.R:418) --------------------------------------------------------------- */
SEXP lp_fit_pi(SEXP sctx, SEXP model, SEXP params, SEXP n_out) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_model m = model_of(model);
    lpb200_params p = params_of(params);
    lpb200_control c = *current_control();
    int n = asInteger(n_out);
    SEXP ll = PROTECT(allocVector(REALSXP, n)), conv = PROTECT(allocVector(INTSXP, n));
    check(ctx, lpb200_fit_pi(ctx, &m, &p, &c, REAL(ll), INTEGER(conv)));
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, params);
    SET_VECTOR_ELT(out, 1, conv);
    SET_VECTOR_ELT(out, 2, ll);
    UNPROTECT(3);
    return out;
}

/* ---- fitModel (fitModel.R:137): whole EM loop in one call -----------------------------------
 * keep_mask: LPB200_KEEP_* bits of the parameter matrices that are warm starts (matMuModelInit, lsmatBatchModelInitMu,
 * matDispModelInit, lsmatBatchModelInitDisp, fitModel.R:147-151); 0 = the library initialises everything */
SEXP lp_fit_model(SEXP sctx, SEXP model, SEXP params, SEXP fit_drop, SEXP n_genes, SEXP keep_mask) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_model m = model_of(model);
    lpb200_params p = params_of(params);
    lpb200_control c = *current_control();
    int G = asInteger(n_genes);
    SEXP em = PROTECT(allocVector(REALSXP, c.max_cycles)), ll = PROTECT(allocVector(REALSXP, G));
    SEXP conv = PROTECT(allocVector(INTSXP, G)), info = PROTECT(allocVector(REALSXP, 3));
    int32_t n_iter = 0, converged = 0;
    double ll_init = 0;
    check(ctx, lpb200_fit_model_from(ctx, &m, &p, asLogical(fit_drop), &c, R_NaN, asInteger(keep_mask), REAL(em), &n_iter,
                                     &converged, &ll_init, REAL(ll), INTEGER(conv)));
    for (int k = 0; k < c.max_cycles; k++)
        if (ISNAN(REAL(em)[k])) REAL(em)[k] = NA_REAL; /* vecEMLogLikModel <- array(NA, ...) fitModel.R:184 */
    REAL(info)[0] = n_iter;
    REAL(info)[1] = converged;
    REAL(info)[2] = ll_init;
    SEXP out = PROTECT(allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, params);
    SET_VECTOR_ELT(out, 1, em);
    SET_VECTOR_ELT(out, 2, ll);
    SET_VECTOR_ELT(out, 3, conv);
    SET_VECTOR_ELT(out, 4, info);
    UNPROTECT(5);
    return out;
}

/* ---- fits B, C, D of fitLPModels (fitWrapperLP.R:152-238) in one concurrent call ------------------
 * models: integer matrix 4 x n_jobs; params: list of n_jobs parameter lists (see params_of) */
SEXP lp_fit_models(SEXP sctx, SEXP models, SEXP params, SEXP n_genes) {
    lpb200_ctx *ctx = ctx_of(sctx);
    int n_jobs = LENGTH(params), G = asInteger(n_genes);
    lpb200_model *m = (lpb200_model *)R_alloc(n_jobs, sizeof(lpb200_model));
    lpb200_params *p = (lpb200_params *)R_alloc(n_jobs, sizeof(lpb200_params));
    for (int j = 0; j < n_jobs; j++) {
        m[j].mu_model = INTEGER(models)[4 * j];
        m[j].disp_model = INTEGER(models)[4 * j + 1];
        m[j].drop_model = INTEGER(models)[4 * j + 2];
        m[j].drop_fit_group = INTEGER(models)[4 * j + 3];
        p[j] = params_of(VECTOR_ELT(params, j));
    }
    lpb200_control c = *current_control();
    SEXP ll_init = PROTECT(allocVector(REALSXP, n_jobs)), em = PROTECT(allocVector(REALSXP, n_jobs));
    SEXP convd = PROTECT(allocVector(INTSXP, n_jobs)), ll = PROTECT(allocVector(REALSXP, (long)n_jobs * G));
    SEXP conv = PROTECT(allocVector(INTSXP, (long)n_jobs * G));
    check(ctx, lpb200_fit_model_multi(ctx, n_jobs, m, p, &c, R_NaN, REAL(ll_init), REAL(em), INTEGER(convd), REAL(ll),
                                      INTEGER(conv), NULL));
    SEXP out = PROTECT(allocVector(VECSXP, 6));
    SET_VECTOR_ELT(out, 0, params);
    SET_VECTOR_ELT(out, 1, ll_init);
    SET_VECTOR_ELT(out, 2, em);
    SET_VECTOR_ELT(out, 3, convd);
    SET_VECTOR_ELT(out, 4, ll);
    SET_VECTOR_ELT(out, 5, conv);
    UNPROTECT(6);
    return out;
}

/* ---- getFitsMean / getFitsDispersion / getFitsDropout / getNormData / getPostDrop (getFits.R) -------- */
SEXP lp_expand_fits(SEXP sctx, SEXP model, SEXP params, SEXP gene_ids, SEXP what, SEXP n_cells) {
    lpb200_ctx *ctx = ctx_of(sctx);
    lpb200_model m = model_of(model);
    lpb200_params p = params_of(params);
    int n = LENGTH(gene_ids);
    SEXP z = PROTECT(allocVector(REALSXP, (long)n * asInteger(n_cells)));
    check(ctx, lpb200_expand_fits(ctx, &m, &p, shift_idx(gene_ids), n, asInteger(what), REAL(z)));
    UNPROTECT(1);
    return z;
}

/* ---- LRT of runDEAnalysis (runHypothesisTests.R:81-86) -------------------------------------- */
SEXP lp_lrt(SEXP ll_full, SEXP ll_red, SEXP ddf) {
    int G = LENGTH(ll_full);
    SEXP p = PROTECT(allocVector(REALSXP, G)), padj = PROTECT(allocVector(REALSXP, G));
    if (lpb200_lrt(REAL(ll_full), REAL(ll_red), asInteger(ddf), G, REAL(p), REAL(padj)) != 0) error("lpb200_lrt failed");
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, p);
    SET_VECTOR_ELT(out, 1, padj);
    UNPROTECT(3);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"lp_set_control", (DL_FUNC)&lp_set_control, 1},
    {"lp_create", (DL_FUNC)&lp_create, 1},
    {"lp_load_counts_csc", (DL_FUNC)&lp_load_counts_csc, 5},
    {"lp_load_counts_dense", (DL_FUNC)&lp_load_counts_dense, 2},
    {"lp_set_design", (DL_FUNC)&lp_set_design, 2},
    {"lp_eval_loglik", (DL_FUNC)&lp_eval_loglik, 4},
    {"lp_fit_mudisp", (DL_FUNC)&lp_fit_mudisp, 4},
    {"lp_fit_pi", (DL_FUNC)&lp_fit_pi, 4},
    {"lp_fit_model", (DL_FUNC)&lp_fit_model, 6},
    {"lp_fit_models", (DL_FUNC)&lp_fit_models, 4},
    {"lp_expand_fits", (DL_FUNC)&lp_expand_fits, 6},
    {"lp_lrt", (DL_FUNC)&lp_lrt, 3},
    {NULL, NULL, 0}};

void R_init_LineagePulse(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
