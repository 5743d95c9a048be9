/* TEST INFRASTRUCTURE.  A minimal runtime behind tests/rstub/Rinternals.h: just enough of R's
 * object model (integer / logical / double vectors and matrices, lists with names, external
 * pointers with finalizers, the routine-registration table) to EXECUTE r-shim/src/vbnmf_shim.c
 * from the Python tests through ctypes.  No garbage collector: objects live until rstub_free_all().
 * Rf_error prints the message and aborts (the tests exercise the success paths). */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <Rinternals.h>
#include <R_ext/Rdynload.h>

struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t length;
    int nrow, ncol;        /* 0, 0 for plain vectors */
    void *data;            /* int / double / SEXP elements, or the external pointer address */
    const char **names;    /* lists made by Rf_mkNamed */
    R_CFinalizer_t fin;
    struct SEXPREC *next;  /* allocation chain */
};

static struct SEXPREC nil_obj = {NILSXP, 0, 0, 0, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
double R_NaReal = 0.0; /* set to a NaN in rstub_init() */
static SEXP chain = NULL;
static const R_CallMethodDef *registered = NULL;
static int dynamic_symbols = -1;

static size_t elt_size(SEXPTYPE t) {
    return t == REALSXP ? sizeof(double) : t == VECSXP ? sizeof(SEXP) : t == RAWSXP ? 1 : sizeof(int);
}

static SEXP new_obj(SEXPTYPE type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type;
    s->length = n;
    if (type != EXTPTRSXP) s->data = calloc((size_t)(n > 0 ? n : 1), elt_size(type));
    if (type == VECSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
    s->next = chain;
    chain = s;
    return s;
}

void rstub_init(void) {
    union { uint64_t u; double d; } na = {0x7FF00000000007A2ull}; /* R's NA_real_ payload 1954 */
    R_NaReal = na.d;
}

void rstub_free_all(void) {
    while (chain) {
        SEXP s = chain;
        chain = s->next;
        if (s->type == EXTPTRSXP && s->fin && s->data) s->fin(s);
        if (s->type != EXTPTRSXP) free(s->data);
        free(s);
    }
}

/* ---- helpers for the Python side --------------------------------------------------------- */
SEXP rstub_real(const double *v, R_xlen_t n, int nrow, int ncol) {
    SEXP s = new_obj(REALSXP, n);
    memcpy(s->data, v, (size_t)n * sizeof(double));
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP rstub_int(const int *v, R_xlen_t n, int logical) {
    SEXP s = new_obj(logical ? LGLSXP : INTSXP, n);
    memcpy(s->data, v, (size_t)n * sizeof(int));
    return s;
}
int rstub_type(SEXP s) { return (int)s->type; }
void *rstub_data(SEXP s) { return s->data; }
const char *rstub_name(SEXP s, int i) { return s->names ? s->names[i] : NULL; }
int rstub_registered(int i, const char **name, int *nargs, void **fun) {
    if (!registered || !registered[i].name) return 0;
    *name = registered[i].name; *nargs = registered[i].numArgs; *fun = (void *)registered[i].fun;
    return 1;
}
int rstub_dynamic_symbols(void) { return dynamic_symbols; }

/* ---- the R API subset ------------------------------------------------------------------------ */
int *INTEGER(SEXP x) { return (int *)x->data; }
int *LOGICAL(SEXP x) { return (int *)x->data; }
double *REAL(SEXP x) { return (double *)x->data; }
Rbyte *RAW(SEXP x) { return (Rbyte *)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x->length; }
int Rf_asInteger(SEXP x) {
    if (x->type == REALSXP) return (int)REAL(x)[0];
    return INTEGER(x)[0];
}
double Rf_asReal(SEXP x) {
    if (x->type == REALSXP) return REAL(x)[0];
    return (double)INTEGER(x)[0];
}
int Rf_ncols(SEXP x) { return x->ncol ? x->ncol : 1; }
int Rf_nrows(SEXP x) { return x->nrow ? x->nrow : (int)x->length; }
Rboolean Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP ? TRUE : FALSE; }
SEXP Rf_ScalarReal(double v) { return rstub_real(&v, 1, 0, 0); }
SEXP Rf_ScalarInteger(int v) { return rstub_int(&v, 1, 0); }
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) { return new_obj(type, n); }
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
    SEXP s = new_obj(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_duplicate(SEXP x) {
    SEXP s = new_obj(x->type, x->length);
    memcpy(s->data, x->data, (size_t)x->length * elt_size(x->type));
    s->nrow = x->nrow; s->ncol = x->ncol; s->names = x->names;
    return s;
}
SEXP Rf_mkNamed(SEXPTYPE type, const char **names) {
    R_xlen_t n = 0;
    while (names[n][0]) n++;
    SEXP s = new_obj(type, n);
    s->names = names;
    return s;
}
SEXP Rf_lengthgets(SEXP x, R_len_t n) {
    SEXP s = new_obj(x->type, n);
    memcpy(s->data, x->data, (size_t)(n < x->length ? n : x->length) * elt_size(x->type));
    return s;
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int n) { (void)n; }
void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Rf_error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    abort();
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = new_obj(EXTPTRSXP, 1);
    s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }
int R_registerRoutines(DllInfo *info, const R_CMethodDef *const c, const R_CallMethodDef *const call,
                       const R_FortranMethodDef *const f, const R_ExternalMethodDef *const e) {
    (void)info; (void)c; (void)f; (void)e;
    registered = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) {
    (void)info;
    dynamic_symbols = (int)value;
    return TRUE;
}
