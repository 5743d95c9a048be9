/* TEST INFRASTRUCTURE.  Stand-in for the subset of R's C API (R_HOME/include/Rinternals.h) that
 * r-shim/src/vbnmf_shim.c uses, with R's own prototypes, so that the shim can be compiled, linked
 * and executed where R is not installed (it is not, in this repository's build environment).
 * tests/rstub/rstub.c implements a minimal runtime behind these declarations. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int R_len_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
#define EXTPTRSXP 22
#define RAWSXP 24
typedef unsigned char Rbyte;

extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal

int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
double *REAL(SEXP x);
Rbyte *RAW(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_ncols(SEXP x);
int Rf_nrows(SEXP x);
Rboolean Rf_isNull(SEXP x);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_duplicate(SEXP x);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_mkNamed(SEXPTYPE type, const char **names);
SEXP Rf_lengthgets(SEXP x, R_len_t n);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#if defined(__GNUC__)
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char *fmt, ...);
#endif

typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

#ifdef __cplusplus
}
#endif
#endif
