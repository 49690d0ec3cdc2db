/* fake_r_runtime.c -- TEST INFRASTRUCTURE: a small functional stand-in for the part of R's C API that
 * r_shim/lp_rshim.c uses (r_shim/fake_R/Rinternals.h): integer / double / list / character vectors with
 * names and dim attributes, external pointers with finalizers, R_alloc, PROTECT, error().  R is not
 * installable in the build image; with this runtime the shim's C code -- argument unpacking, 1-based index
 * shifting, result packing -- is linked and RUN against liblpb200.so on the GPU
 * (tests/test_gpu_parity.py::test_r_shim_runs_on_the_gpu), instead of being syntax-checked only.
 * Memory is never reclaimed (a test process): no garbage collector, PROTECT / UNPROTECT are no-ops. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#define CHARSXP 9
#define STRSXP 16
#define EXTPTRSXP 22
#define NILSXP 0
#define SYMSXP 1

struct SEXPREC {
    int type;
    long length;
    void *data;          /* int[], double[], SEXP[], char[] */
    SEXP names, dim;     /* the two attributes the shim reads */
    void *ext;           /* external pointer payload */
    R_CFinalizer_t fin;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC names_sym = {SYMSXP, 0, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC dim_sym = {SYMSXP, 0, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
double R_NaN, R_NaReal;

static jmp_buf *err_jmp = NULL;
static char err_msg[1024];
const char *fake_r_last_error(void) { return err_msg; }
void fake_r_set_jmp(jmp_buf *j) { err_jmp = j; }

static void __attribute__((constructor)) init_consts(void) {
    R_NaN = NAN;
    union { double d; unsigned long long u; } na;
    na.u = 0x7FF00000000007A2ull;   /* R's NA_real_: a NaN with payload 1954 */
    R_NaReal = na.d;
}

static SEXP new_rec(int type, long n) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type;
    s->length = n;
    s->names = s->dim = R_NilValue;
    return s;
}

SEXP allocVector(unsigned int type, long n) {
    SEXP s = new_rec((int)type, n);
    size_t el = type == INTSXP ? sizeof(int) : type == REALSXP ? sizeof(double) : sizeof(SEXP);
    s->data = calloc((size_t)(n > 0 ? n : 1), el);
    if (type == VECSXP || type == STRSXP)
        for (long k = 0; k < n; k++) ((SEXP *)s->data)[k] = R_NilValue;
    return s;
}
SEXP mkChar(const char *c) {
    SEXP s = new_rec(CHARSXP, (long)strlen(c));
    s->data = strdup(c);
    return s;
}
SEXP getAttrib(SEXP x, SEXP which) {
    if (x == R_NilValue) return R_NilValue;
    if (which == R_NamesSymbol) return x->names;
    if (which == R_DimSymbol) return x->dim;
    return R_NilValue;
}
SEXP setAttrib(SEXP x, SEXP which, SEXP v) {
    if (which == R_NamesSymbol) x->names = v;
    else if (which == R_DimSymbol) x->dim = v;
    return v;
}
int LENGTH(SEXP x) { return (int)x->length; }
SEXP STRING_ELT(SEXP x, int i) { return ((SEXP *)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, int i) { return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, int i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
void SET_STRING_ELT(SEXP x, int i, SEXP v) { ((SEXP *)x->data)[i] = v; }
const char *CHAR(SEXP x) { return (const char *)x->data; }
double *REAL(SEXP x) {
    if (x->type != REALSXP) error("REAL() can only be applied to a 'numeric', not a type %d object", x->type);
    return (double *)x->data;
}
int *INTEGER(SEXP x) {
    if (x->type != INTSXP) error("INTEGER() can only be applied to a 'integer', not a type %d object", x->type);
    return (int *)x->data;
}
int asInteger(SEXP x) {
    if (x == R_NilValue || x->length < 1) return (int)0x80000000;   /* NA_integer_ */
    if (x->type == INTSXP) return ((int *)x->data)[0];
    if (x->type == REALSXP) return (int)((double *)x->data)[0];
    return (int)0x80000000;
}
int asLogical(SEXP x) { return asInteger(x) != 0; }
SEXP PROTECT(SEXP x) { return x; }
void UNPROTECT(int n) { (void)n; }
char *R_alloc(size_t n, int size) { return (char *)calloc(n ? n : 1, (size_t)size); }
void error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof err_msg, fmt, ap);
    va_end(ap);
    if (err_jmp) longjmp(*err_jmp, 1);
    fprintf(stderr, "fake R error: %s\n", err_msg);
    abort();
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = new_rec(EXTPTRSXP, 1);
    s->ext = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->ext; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fin, Rboolean onexit) { (void)onexit; s->fin = fin; }
void fake_r_finalize(SEXP s) { if (s->fin) s->fin(s); }   /* what gc() would do */
int R_registerRoutines(DllInfo *d, const void *a, const R_CallMethodDef *c, const void *b, const void *e) {
    (void)d; (void)a; (void)c; (void)b; (void)e;
    return 1;
}
int R_useDynamicSymbols(DllInfo *d, int v) { (void)d; (void)v; return 0; }
