/* rshim_driver.c -- TEST INFRASTRUCTURE: plays the part of the R wrappers (r_shim/R/lpb200_wrappers.R) against
 * r_shim/lp_rshim.c under the stand-in R runtime (fake_r_runtime.c): builds the SEXP arguments the wrappers build
 * (1-based index vectors, named lists, dgCMatrix slots), calls the .Call entry points in the order of
 * fitLPModels + runDEAnalysis, and copies the results out for comparison with the ctypes path. */
#include <setjmp.h>
#include <string.h>
#include <Rinternals.h>

SEXP lp_create(SEXP), lp_load_counts_csc(SEXP, SEXP, SEXP, SEXP, SEXP), lp_set_design(SEXP, SEXP);
SEXP lp_fit_model(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP), lp_fit_models(SEXP, SEXP, SEXP, SEXP), lp_eval_loglik(SEXP, SEXP, SEXP, SEXP);
SEXP lp_lrt(SEXP, SEXP, SEXP), lp_fit_pi(SEXP, SEXP, SEXP, SEXP), lp_fit_mudisp(SEXP, SEXP, SEXP, SEXP);
SEXP lp_expand_fits(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP), lp_set_control(SEXP);
void fake_r_set_jmp(jmp_buf *);
const char *fake_r_last_error(void);
void fake_r_finalize(SEXP);

static SEXP ints(const int *v, int n) {
    SEXP s = allocVector(INTSXP, n);
    memcpy(INTEGER(s), v, sizeof(int) * n);
    return s;
}
static SEXP int1(int v) { return ints(&v, 1); }
static SEXP reals(const double *v, long n) {
    SEXP s = allocVector(REALSXP, n);
    if (v) memcpy(REAL(s), v, sizeof(double) * n);
    return s;
}
static SEXP matrix_of(const double *v, int nr, int nc) {
    SEXP s = reals(v, (long)nr * nc);
    int d[2] = {nr, nc};
    setAttrib(s, R_DimSymbol, ints(d, 2));
    return s;
}
static SEXP named_list(int n, const char **names, SEXP *vals) {
    SEXP l = allocVector(VECSXP, n), nm = allocVector(16 /* STRSXP */, n);
    for (int k = 0; k < n; k++) {
        SET_VECTOR_ELT(l, k, vals[k]);
        SET_STRING_ELT(nm, k, mkChar(names[k]));
    }
    setAttrib(l, R_NamesSymbol, nm);
    return l;
}
/* lpParams(): list(matMuModel, matBatchMu, matDispModel, matBatchDisp, matDropoutLinModel) */
static SEXP params_list(int G, int N, int p_mu, const double *drop /* N x 1 or NULL */) {
    const char *nm[5] = {"matMuModel", "matBatchMu", "matDispModel", "matBatchDisp", "matDropoutLinModel"};
    SEXP v[5] = {matrix_of(NULL, G, p_mu), R_NilValue, matrix_of(NULL, G, 1), R_NilValue, drop ? matrix_of(drop, N, 1) : R_NilValue};
    return named_list(5, nm, v);
}
static SEXP list_get(SEXP list, const char *name) {
    SEXP names = getAttrib(list, R_NamesSymbol);
    for (int i = 0; i < LENGTH(list); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}

/* impulse vs constant, ZINB, logistic drop-out: fit A (EM), fits B / C / D concurrently, LRT.
 * Inputs: dgCMatrix slots (p [N + 1], i, x), pseudotime [N], impulse basis [N x 4] column-major, device ids.
 * Outputs: drop [N], mu_h1 [G x 7], disp_h1 [G], mu_h0 [G], ll_h1 / ll_h0 / p / padj [G], em_ll [20], info [3]. */
int rs_run_impulse_pipeline(int G, int N, const int *cp, const int *ci, const double *cx, const double *t, const double *basis,
                            const int *devices, int n_dev, double *drop, double *mu_h1, double *disp_h1, double *mu_h0,
                            double *ll_h1, double *ll_h0, double *p, double *padj, double *em_ll, double *info, char *errbuf) {
    jmp_buf jb;
    fake_r_set_jmp(&jb);
    if (setjmp(jb)) {
        strncpy(errbuf, fake_r_last_error(), 511);
        fake_r_set_jmp(NULL);
        return 1;
    }
    SEXP ctx = lp_create(ints(devices, n_dev));
    int dim[2] = {G, N};
    lp_load_counts_csc(ctx, ints(dim, 2), ints(cp, N + 1), ints(ci, cp[N]), reals(cx, cp[N]));
    /* lpDesign(): vecNormConst = 1, pseudotime, impulse basis; everything else NULL */
    SEXP sf = allocVector(REALSXP, N);
    for (int j = 0; j < N; j++) REAL(sf)[j] = 1.0;
    const char *dn[12] = {"n_genes", "n_cells", "vecNormConst", "vecContinuousCovar", "vecidxGroups", "scaNGroups",
                          "matSplineBasisMu", "matSplineBasisDisp", "matImpulseBasis", "vecNBatchesMu", "vecNBatchesDisp",
                          "matPiConstPredictors"};
    SEXP dv[12] = {int1(G), int1(N), sf, reals(t, N), R_NilValue, int1(0), R_NilValue, R_NilValue, matrix_of(basis, N, 4),
                   R_NilValue, R_NilValue, R_NilValue};
    lp_set_design(ctx, named_list(12, dn, dv));
    /* model codes as lpModel(): c(mu, disp, drop, group) */
    int mA[4] = {0, 0, 1, 0}, mB[4] = {1, 0, 1, 0}, mC[4] = {0, 0, 0, 0}, mD[4] = {1, 0, 0, 0};
    SEXP pA = params_list(G, N, 1, NULL);
    SET_VECTOR_ELT(pA, 4, matrix_of(NULL, N, 1));   /* matDropoutLinModel to be estimated */
    SEXP rA = lp_fit_model(ctx, ints(mA, 4), pA, int1(1), int1(G), int1(0));
    SEXP pAo = VECTOR_ELT(rA, 0);
    memcpy(drop, REAL(list_get(pAo, "matDropoutLinModel")), sizeof(double) * N);
    memcpy(mu_h0, REAL(list_get(pAo, "matMuModel")), sizeof(double) * G);
    memcpy(em_ll, REAL(VECTOR_ELT(rA, 1)), sizeof(double) * 20);
    memcpy(info, REAL(VECTOR_ELT(rA, 4)), sizeof(double) * 3);
    /* fits B, C, D */
    int models[12];
    memcpy(models, mB, sizeof mB);
    memcpy(models + 4, mC, sizeof mC);
    memcpy(models + 8, mD, sizeof mD);
    SEXP plist = allocVector(VECSXP, 3);
    SET_VECTOR_ELT(plist, 0, params_list(G, N, 7, drop));
    SET_VECTOR_ELT(plist, 1, params_list(G, N, 1, NULL));
    SET_VECTOR_ELT(plist, 2, params_list(G, N, 7, NULL));
    SEXP rM = lp_fit_models(ctx, ints(models, 12), plist, int1(G));
    SEXP pB = VECTOR_ELT(VECTOR_ELT(rM, 0), 0);
    memcpy(mu_h1, REAL(list_get(pB, "matMuModel")), sizeof(double) * G * 7);
    memcpy(disp_h1, REAL(list_get(pB, "matDispModel")), sizeof(double) * G);
    /* runDEAnalysis: evalLogLikMatrix of H1 and H0, LRT with 6 degrees of freedom */
    SEXP llf = lp_eval_loglik(ctx, ints(mB, 4), pB, int1(G));
    SEXP llr = lp_eval_loglik(ctx, ints(mA, 4), pAo, int1(G));
    memcpy(ll_h1, REAL(llf), sizeof(double) * G);
    memcpy(ll_h0, REAL(llr), sizeof(double) * G);
    SEXP lr = lp_lrt(llf, llr, int1(6));
    memcpy(p, REAL(VECTOR_ELT(lr, 0)), sizeof(double) * G);
    memcpy(padj, REAL(VECTOR_ELT(lr, 1)), sizeof(double) * G);
    fake_r_finalize(ctx);   /* what R's garbage collector does with the external pointer */
    fake_r_set_jmp(NULL);
    return 0;
}

/* error path: .Call on a released context must raise an R error, not crash */
int rs_released_context_raises(char *errbuf) {
    jmp_buf jb;
    fake_r_set_jmp(&jb);
    if (setjmp(jb)) {
        strncpy(errbuf, fake_r_last_error(), 511);
        fake_r_set_jmp(NULL);
        return 1;
    }
    int dev = 0;
    SEXP ctx = lp_create(ints(&dev, 1));
    fake_r_finalize(ctx);
    int m[4] = {0, 0, 0, 0};
    lp_eval_loglik(ctx, ints(m, 4), params_list(4, 4, 1, NULL), int1(4));
    fake_r_set_jmp(NULL);
    return 0;
}

/* lp_set_control: NULL and NA entries keep fitModel.R:164-175's constants, values are taken, nonsense raises an R error
 * and leaves the defaults.  Returns 0 when all of that holds (no GPU needed). */
int rs_set_control_checks(char *errbuf) {
    jmp_buf jb;
    fake_r_set_jmp(&jb);
    if (setjmp(jb)) {
        strncpy(errbuf, fake_r_last_error(), 511);
        fake_r_set_jmp(NULL);
        return 100;
    }
    lp_set_control(R_NilValue);
    double ok[5] = {500, 1e-6, R_NaReal, R_NaReal, 5};
    lp_set_control(reals(ok, 5));
    double shorter[2] = {R_NaReal, 1e-7};
    lp_set_control(reals(shorter, 2));
    fake_r_set_jmp(NULL);
    jmp_buf jb2;
    fake_r_set_jmp(&jb2);
    if (setjmp(jb2)) {
        strncpy(errbuf, fake_r_last_error(), 511);
        fake_r_set_jmp(NULL);
        lp_set_control(R_NilValue);
        return 0;   /* the bad vector raised */
    }
    double bad[5] = {1000, -1.0, 10000, 1e-8, 0};
    lp_set_control(reals(bad, 5));
    fake_r_set_jmp(NULL);
    return 1;       /* no error: wrong */
}
