/* TEST INFRASTRUCTURE: a minimal stand-in for R's C API (R.h / Rinternals.h), just the entry points
 * bart-int_b200/R/r_shim.c uses, so the shim can be compiled against include/bartint.h and driven from the tests
 * on a machine without R.  Semantics follow "Writing R Extensions" for these calls (column-major REAL/INTEGER
 * payloads, dim attribute as an INTSXP of length 2, CHARSXP strings, external pointers with C finalizers,
 * Rf_error never returns).  PROTECT is a no-op: the mock never collects. */
#ifndef BARTINT_R_MOCK_RINTERNALS_H
#define BARTINT_R_MOCK_RINTERNALS_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
typedef void (*R_CFinalizer_t)(SEXP);

#define NILSXP 0
#define CHARSXP 9
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22

extern SEXP R_NilValue;
extern SEXP R_DimSymbol;

SEXP Rf_getAttrib(SEXP x, SEXP name);
int* INTEGER(SEXP x);
double* REAL(SEXP x);
const char* CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
R_xlen_t XLENGTH(SEXP x);
char* R_alloc(size_t n, int size);
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
Rboolean Rf_isNull(SEXP x);
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
SEXP Rf_ScalarInteger(int v);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#if defined(__GNUC__)
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char* fmt, ...);
#endif
#define PROTECT(s) (s)
#define UNPROTECT(n) ((void)(n))

#ifdef __cplusplus
}
#endif
#endif
