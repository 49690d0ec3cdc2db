/* Minimal stand-in for R's Rinternals.h: declares ONLY what r_shim/lp_rshim.c uses (plus the constructors the
 * test driver needs), so that the shim can be compiled -- and, linked with tests/csrc/fake_r_runtime.c, RUN --
 * in an image without R. */
#ifndef FAKE_RINTERNALS_H
#define FAKE_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaN, R_NaReal;
#define NA_REAL R_NaReal
int R_IsNaN(double), R_IsNA(double);
#define ISNAN(x) ((x) != (x))
SEXP getAttrib(SEXP, SEXP);
int LENGTH(SEXP);
SEXP STRING_ELT(SEXP, int);
SEXP VECTOR_ELT(SEXP, int);
SEXP SET_VECTOR_ELT(SEXP, int, SEXP);
const char *CHAR(SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
int asInteger(SEXP);
int asLogical(SEXP);
SEXP allocVector(unsigned int, long);
SEXP mkChar(const char *);
SEXP setAttrib(SEXP, SEXP, SEXP);
void SET_STRING_ELT(SEXP, int, SEXP);
SEXP PROTECT(SEXP);
void UNPROTECT(int);
char *R_alloc(size_t, int);
void error(const char *, ...);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
