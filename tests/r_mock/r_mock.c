/* TEST INFRASTRUCTURE: implementation of the mock R C API declared in Rinternals.h, plus the helpers the Python
 * tests use to build R values and to .Call the shim's entry points with R's error semantics (Rf_error longjmps
 * back to the caller, the message is kept). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"

struct SEXPREC {
    unsigned int type;
    R_xlen_t length;
    void* data;            /* int*, double*, SEXP* (STRSXP / VECSXP), char* (CHARSXP), void* (EXTPTRSXP) */
    SEXP dim;              /* the only attribute the shim reads */
    R_CFinalizer_t finalizer;
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL};
static struct SEXPREC dim_symbol_rec = {NILSXP, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_DimSymbol = &dim_symbol_rec;

static jmp_buf* error_target = NULL;
static char error_message[1024] = "";

/* everything the mock allocates lives until mock_reset(): R_alloc memory and SEXPs alike */
static void** arena = NULL;
static size_t arena_n = 0, arena_cap = 0;
static void* arena_alloc(size_t bytes) {
    void* p = calloc(1, bytes ? bytes : 1);
    if (!p) abort();
    if (arena_n == arena_cap) {
        arena_cap = arena_cap ? 2 * arena_cap : 256;
        arena = (void**)realloc(arena, arena_cap * sizeof(void*));
        if (!arena) abort();
    }
    arena[arena_n++] = p;
    return p;
}

static SEXP new_sexp(unsigned int type, R_xlen_t n, size_t elt) {
    SEXP s = (SEXP)arena_alloc(sizeof(struct SEXPREC));
    s->type = type;
    s->length = n;
    s->data = elt ? arena_alloc((size_t)n * elt) : NULL;
    s->dim = R_NilValue;
    return s;
}

void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_message, sizeof(error_message), fmt, ap);
    va_end(ap);
    if (error_target) longjmp(*error_target, 1);
    fprintf(stderr, "Rf_error outside mock_call: %s\n", error_message);
    abort();
}

static void type_check(SEXP x, unsigned int type, const char* who) {
    if (x == NULL || x->type != type) Rf_error("mock R: %s applied to an object of type %u", who, x ? x->type : 999u);
}

SEXP Rf_getAttrib(SEXP x, SEXP name) { return (name == R_DimSymbol && x) ? x->dim : R_NilValue; }
int* INTEGER(SEXP x) { type_check(x, INTSXP, "INTEGER"); return (int*)x->data; }
double* REAL(SEXP x) { type_check(x, REALSXP, "REAL"); return (double*)x->data; }
const char* CHAR(SEXP x) { type_check(x, CHARSXP, "CHAR"); return (const char*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) {
    type_check(x, STRSXP, "STRING_ELT");
    if (i < 0 || i >= x->length) Rf_error("mock R: STRING_ELT index %ld out of range", (long)i);
    return ((SEXP*)x->data)[i];
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    type_check(x, VECSXP, "VECTOR_ELT");
    if (i < 0 || i >= x->length) Rf_error("mock R: VECTOR_ELT index %ld out of range", (long)i);
    return ((SEXP*)x->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    type_check(x, VECSXP, "SET_VECTOR_ELT");
    if (i < 0 || i >= x->length) Rf_error("mock R: SET_VECTOR_ELT index %ld out of range", (long)i);
    ((SEXP*)x->data)[i] = v;
    return v;
}
R_xlen_t XLENGTH(SEXP x) { return x ? x->length : 0; }
char* R_alloc(size_t n, int size) { return (char*)arena_alloc(n * (size_t)size); }
double Rf_asReal(SEXP x) {
    if (x && x->length >= 1 && x->type == REALSXP) return ((double*)x->data)[0];
    if (x && x->length >= 1 && x->type == INTSXP) return (double)((int*)x->data)[0];
    Rf_error("mock R: asReal of a non-numeric value");
}
int Rf_asInteger(SEXP x) {
    if (x && x->length >= 1 && x->type == INTSXP) return ((int*)x->data)[0];
    if (x && x->length >= 1 && x->type == REALSXP) return (int)((double*)x->data)[0];
    Rf_error("mock R: asInteger of a non-numeric value");
}
Rboolean Rf_isNull(SEXP x) { return (x == NULL || x->type == NILSXP) ? TRUE : FALSE; }
SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
    switch (type) {
        case INTSXP: return new_sexp(type, n, sizeof(int));
        case REALSXP: return new_sexp(type, n, sizeof(double));
        case STRSXP:
        case VECSXP: {
            SEXP s = new_sexp(type, n, sizeof(SEXP));
            for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
            return s;
        }
        default: Rf_error("mock R: allocVector of type %u", type);
    }
}
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->dim = Rf_allocVector(INTSXP, 2);
    INTEGER(s->dim)[0] = nrow;
    INTEGER(s->dim)[1] = ncol;
    return s;
}

SEXP Rf_ScalarInteger(int v) {
    SEXP s = Rf_allocVector(INTSXP, 1);
    INTEGER(s)[0] = v;
    return s;
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = new_sexp(EXTPTRSXP, 1, 0);
    s->data = p;
    return s;
}
void* R_ExternalPtrAddr(SEXP s) { type_check(s, EXTPTRSXP, "R_ExternalPtrAddr"); return s->data; }
void R_ClearExternalPtr(SEXP s) { type_check(s, EXTPTRSXP, "R_ClearExternalPtr"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
    (void)onexit;
    type_check(s, EXTPTRSXP, "R_RegisterCFinalizerEx");
    s->finalizer = fun;
}

/* ---- helpers for the Python side ----------------------------------------------------------------------------- */
SEXP mock_nil(void) { return R_NilValue; }
SEXP mock_real(const double* v, R_xlen_t n) {
    SEXP s = Rf_allocVector(REALSXP, n);
    if (n) memcpy(s->data, v, (size_t)n * sizeof(double));
    return s;
}
SEXP mock_int(const int* v, R_xlen_t n) {
    SEXP s = Rf_allocVector(INTSXP, n);
    if (n) memcpy(s->data, v, (size_t)n * sizeof(int));
    return s;
}
void mock_set_dim(SEXP s, const int* dims, int nd) { s->dim = mock_int(dims, nd); }
SEXP mock_strings(const char* const* v, R_xlen_t n) {
    SEXP s = Rf_allocVector(STRSXP, n);
    for (R_xlen_t i = 0; i < n; ++i) {
        SEXP c = new_sexp(CHARSXP, (R_xlen_t)strlen(v[i]), 0);
        c->data = arena_alloc(strlen(v[i]) + 1);
        strcpy((char*)c->data, v[i]);
        ((SEXP*)s->data)[i] = c;
    }
    return s;
}
SEXP mock_list(R_xlen_t n) { return Rf_allocVector(VECSXP, n); }
unsigned int mock_type(SEXP s) { return s ? s->type : 999u; }
R_xlen_t mock_length(SEXP s) { return s ? s->length : 0; }
const double* mock_real_ptr(SEXP s) { return (s && s->type == REALSXP) ? (const double*)s->data : NULL; }
const int* mock_dim_ptr(SEXP s) { return (s && s->dim != R_NilValue) ? (const int*)s->dim->data : NULL; }
const char* mock_last_error(void) { return error_message; }

/* gc(): run the finalizer of an external pointer as R would when the handle becomes unreachable */
void mock_finalize(SEXP s) {
    if (s && s->type == EXTPTRSXP && s->finalizer) { s->finalizer(s); s->finalizer = NULL; }
}

void mock_reset(void) {
    for (size_t i = 0; i < arena_n; ++i) free(arena[i]);
    arena_n = 0;
}

/* .Call with up to 11 arguments: returns the value, or NULL after Rf_error (message in mock_last_error()) */
typedef SEXP (*dotcall_t)();
SEXP mock_call(dotcall_t fn, int nargs, SEXP* a) {
    jmp_buf here;
    jmp_buf* outer = error_target;
    volatile SEXP result = NULL;
    error_message[0] = '\0';
    error_target = &here;
    if (setjmp(here) == 0) {
        switch (nargs) {
            case 1: result = fn(a[0]); break;
            case 2: result = fn(a[0], a[1]); break;
            case 3: result = fn(a[0], a[1], a[2]); break;
            case 5: result = fn(a[0], a[1], a[2], a[3], a[4]); break;
            case 4: result = fn(a[0], a[1], a[2], a[3]); break;
            case 6: result = fn(a[0], a[1], a[2], a[3], a[4], a[5]); break;
            case 11: result = fn(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]); break;
            default: snprintf(error_message, sizeof(error_message), "mock_call: %d arguments not supported", nargs);
        }
    } else {
        result = NULL;
    }
    error_target = outer;
    return result;
}
