/*
 * r_shim.c -- .Call glue between R and libbartint_cuda.so (include/bartint.h).
 *
 * SOURCE ONLY: this image has no R (no Rinternals.h / libR.so), so the shim is not built or tested here;
 * everything it calls is exercised through the same C ABI by the ctypes tests.  On an R box:
 *     R CMD SHLIB r_shim.c -I<repo>/include -L<repo>/bart-int_b200 -lbartint_cuda -o bartint_r.so
 *
 * R objects are read in place (column-major doubles, CHARSXP strings) -- no conversion, nothing retained.
 * Every non-zero status becomes an R error carrying bartint_last_error().
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>

#include "bartint.h"

static void check(int status) {
    if (status != BARTINT_OK) Rf_error("bartint: %s", bartint_last_error());
}

static void forest_finalizer(SEXP ptr) {
    bartint_forest* f = (bartint_forest*)R_ExternalPtrAddr(ptr);
    if (f) { bartint_forest_destroy(f); R_ClearExternalPtr(ptr); }
}

static bartint_forest* forest_of(SEXP ptr) {
    bartint_forest* f = (bartint_forest*)R_ExternalPtrAddr(ptr);
    if (!f) Rf_error("bartint: forest handle was already released");
    return f;
}

/* .Call("bartint_R_forest", savedTrees (chr m x S), savedTreeFits (dbl n x m x S), x (dbl n x d),
 *       cutPoints (list of d numeric), ymin, ymax)  -> external pointer
 * Replaces getTree() for every (tree, draw): src/BARTInt.R:92-133. */
SEXP bartint_R_forest(SEXP trees, SEXP fits, SEXP x, SEXP cuts, SEXP ymin, SEXP ymax) {
    SEXP dim = Rf_getAttrib(trees, R_DimSymbol);
    const int m = INTEGER(dim)[0], S = INTEGER(dim)[1];
    SEXP xdim = Rf_getAttrib(x, R_DimSymbol);
    const int64_t n = INTEGER(xdim)[0];
    const int d = INTEGER(xdim)[1];
    const char** strs = (const char**)R_alloc((size_t)m * S, sizeof(char*));
    for (R_xlen_t k = 0; k < (R_xlen_t)m * S; ++k) strs[k] = CHAR(STRING_ELT(trees, k));   /* [tree, sample] col-major */
    int32_t* off = (int32_t*)R_alloc((size_t)d + 1, sizeof(int32_t));
    off[0] = 0;
    for (int j = 0; j < d; ++j) off[j + 1] = off[j] + (int32_t)XLENGTH(VECTOR_ELT(cuts, j));
    double* vals = (double*)R_alloc((size_t)off[d] + 1, sizeof(double));
    for (int j = 0; j < d; ++j) {
        const double* c = REAL(VECTOR_ELT(cuts, j));
        for (int k = 0; k < off[j + 1] - off[j]; ++k) vals[off[j] + k] = c[k];
    }
    bartint_forest* f = NULL;
    check(bartint_forest_from_dbarts098(strs, REAL(fits), REAL(x), n, d, off, vals, Rf_asReal(ymin), Rf_asReal(ymax),
                                        m, S, &f));
    SEXP ptr = PROTECT(R_MakeExternalPtr(f, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, forest_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call("bartint_R_forest_wbart", fit$treedraws$trees (chr 1), fit$treedraws$cutpoints (list of p numeric),
 *       fit$offset, fit$mu, fit$binaryOffset)  -> external pointer.   For fits made with the BART package instead of
 * dbarts (the reference's README.md:202 invites swapping the BART package); routing follows tree::bn (left iff
 * x < cut).  The centring constant predict() adds is fit$mu for wbart, fit$offset for gbart and fit$binaryOffset for
 * pbart [BART-mem]: the first of the three that is not NULL is used, and it is an error if all are -- silently
 * assuming 0 would shift every prediction, draw mean and integral by mean(y.train). */
SEXP bartint_R_forest_wbart(SEXP trees, SEXP cuts, SEXP offset, SEXP mu, SEXP binary_offset) {
    const int d = (int)XLENGTH(cuts);
    int32_t* off = (int32_t*)R_alloc((size_t)d + 1, sizeof(int32_t));
    off[0] = 0;
    for (int j = 0; j < d; ++j) off[j + 1] = off[j] + (int32_t)XLENGTH(VECTOR_ELT(cuts, j));
    double* vals = (double*)R_alloc((size_t)off[d] + 1, sizeof(double));
    for (int j = 0; j < d; ++j) {
        const double* c = REAL(VECTOR_ELT(cuts, j));
        for (int k = 0; k < off[j + 1] - off[j]; ++k) vals[off[j] + k] = c[k];
    }
    SEXP centre = !Rf_isNull(offset) ? offset : !Rf_isNull(mu) ? mu : binary_offset;
    if (Rf_isNull(centre)) Rf_error("bartint: the fit has none of $offset, $mu, $binaryOffset: cannot tell what predict() adds to the tree sums");
    bartint_forest* f = NULL;
    check(bartint_forest_from_wbart(CHAR(STRING_ELT(trees, 0)), 0, d, off, vals, Rf_asReal(centre), &f));
    SEXP ptr = PROTECT(R_MakeExternalPtr(f, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, forest_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call("bartint_R_forest_values", S, m, d, tree_node_offset (dbl S*m+1), var (int, -1 = leaf, 0-based), value (dbl),
 *       left, right (int or NULL: pre-order), route, leaf_scale, offset)  -> external pointer.
 * Trees exported with split VALUES: dbarts >= 0.9-9 fit$getTrees(), bartMachine::extract_raw_node_data (README.md:
 * 102-106, :202); see bartint_forest_from_value_splits. */
SEXP bartint_R_forest_values(SEXP S, SEXP m, SEXP d, SEXP offs, SEXP var, SEXP value, SEXP left, SEXP right, SEXP route,
                             SEXP leaf_scale, SEXP offset) {
    const R_xlen_t nt = XLENGTH(offs);
    int64_t* off = (int64_t*)R_alloc((size_t)nt, sizeof(int64_t));
    for (R_xlen_t k = 0; k < nt; ++k) off[k] = (int64_t)REAL(offs)[k];
    if (nt != (R_xlen_t)Rf_asInteger(S) * Rf_asInteger(m) + 1 || off[nt - 1] != (int64_t)XLENGTH(var) || XLENGTH(var) != XLENGTH(value))
        Rf_error("bartint: tree offsets, var and value do not describe S * m trees");
    const int has_children = !Rf_isNull(left) && !Rf_isNull(right);
    bartint_forest* f = NULL;
    check(bartint_forest_from_value_splits(Rf_asInteger(S), Rf_asInteger(m), Rf_asInteger(d), off, INTEGER(var), REAL(value),
                                           has_children ? INTEGER(left) : NULL, has_children ? INTEGER(right) : NULL,
                                           Rf_asInteger(route), Rf_asReal(leaf_scale), Rf_asReal(offset), &f));
    SEXP ptr = PROTECT(R_MakeExternalPtr(f, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, forest_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* bartint_init(n_gpus): open the device group (all GPUs behind this one R process).  Returns the group size. */
SEXP bartint_R_init(SEXP n_gpus) {
    check(bartint_init(Rf_asInteger(n_gpus)));
    return Rf_ScalarInteger(bartint_group_size());
}

/* release a forest handle now (the finalizer then finds it cleared) */
SEXP bartint_R_release(SEXP ptr) {
    forest_finalizer(ptr);
    return R_NilValue;
}

/* order(values, decreasing = TRUE)[1:k], 0-based (-1 once the non-NaN values are exhausted) */
SEXP bartint_R_topk(SEXP values, SEXP k) {
    const int kk = Rf_asInteger(k);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, kk > 0 ? kk : 0));
    if (kk > 0) {
        int64_t* idx = (int64_t*)R_alloc((size_t)kk, sizeof(int64_t));
        check(bartint_topk(REAL(values), (int64_t)XLENGTH(values), kk, idx, NULL));
        for (int r = 0; r < kk; ++r) INTEGER(out)[r] = (int)idx[r];
    }
    UNPROTECT(1);
    return out;
}

/* sampleIntegrals(model, dim, measure): src/BARTInt.R:204-223.  n_draws_used <= 0 -> all draws. */
SEXP bartint_R_exact_integrals(SEXP ptr, SEXP measure, SEXP n_draws_used) {
    bartint_forest* f = forest_of(ptr);
    int64_t info[12];
    check(bartint_forest_info(f, info));
    int n = Rf_asInteger(n_draws_used);
    if (n <= 0 || n > info[0]) n = (int)info[0];
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    check(bartint_exact_integrals(f, Rf_asInteger(measure), NULL, NULL, n, REAL(out)));
    UNPROTECT(1);
    return out;
}

/* what: bit 1 = rowMeans, bit 2 = colVars * weights, bit 4 = full S x N matrix (predict).
 * Replaces predict(model, X) / colVars / rowMeans: BARTInt.R:312-314, bartMean.R:47-50,74-82. */
SEXP bartint_R_eval(SEXP ptr, SEXP x, SEXP weights, SEXP what) {
    bartint_forest* f = forest_of(ptr);
    int64_t info[12];
    check(bartint_forest_info(f, info));
    SEXP xdim = Rf_getAttrib(x, R_DimSymbol);
    const int64_t N = INTEGER(xdim)[0];
    const int d = INTEGER(xdim)[1], S = (int)info[0], w = Rf_asInteger(what);
    SEXP rm = PROTECT((w & 1) ? Rf_allocVector(REALSXP, S) : R_NilValue);
    SEXP cv = PROTECT((w & 2) ? Rf_allocVector(REALSXP, N) : R_NilValue);
    SEXP full = PROTECT((w & 4) ? Rf_allocMatrix(REALSXP, S, (int)N) : R_NilValue);
    const int has_w = !Rf_isNull(weights);
    check(bartint_eval(f, REAL(x), N, d, 0u, has_w ? REAL(weights) : NULL, has_w ? XLENGTH(weights) : 0,
                       (w & 1) ? REAL(rm) : NULL, NULL, (w & 2) ? REAL(cv) : NULL, (w & 4) ? REAL(full) : NULL));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, rm); SET_VECTOR_ELT(out, 1, cv); SET_VECTOR_ELT(out, 2, full);
    UNPROTECT(4);
    return out;
}

/* .Call("bartint_R_design_step", savedTrees, savedTreeFits, x, cutPoints, ymin, ymax, candidates (dbl C x d or NULL),
 *       weights (dbl, length 1 or C, or NULL), X (dbl N x d or NULL), measure, n_draws_integral)
 * -> list(integrals, integral_mean_sd, cand_var, draw_mean, draw_mean_var).   One library call for the body of the
 * sequential loop (BARTInt.R:259-262 + 312-314; with X also bartMean.R:74-82): the export of the saved trees is
 * pipelined against the evaluation kernels, so no forest handle is kept. */
SEXP bartint_R_design_step(SEXP trees, SEXP fits, SEXP x, SEXP cuts, SEXP ymin, SEXP ymax, SEXP cand, SEXP weights,
                           SEXP X, SEXP measure, SEXP n_draws_integral) {
    SEXP dim = Rf_getAttrib(trees, R_DimSymbol);
    const int m = INTEGER(dim)[0], S = INTEGER(dim)[1];
    SEXP xdim = Rf_getAttrib(x, R_DimSymbol);
    const int64_t n = INTEGER(xdim)[0];
    const int d = INTEGER(xdim)[1];
    const char** strs = (const char**)R_alloc((size_t)m * S, sizeof(char*));
    for (R_xlen_t k = 0; k < (R_xlen_t)m * S; ++k) strs[k] = CHAR(STRING_ELT(trees, k));
    int32_t* off = (int32_t*)R_alloc((size_t)d + 1, sizeof(int32_t));
    off[0] = 0;
    for (int j = 0; j < d; ++j) off[j + 1] = off[j] + (int32_t)XLENGTH(VECTOR_ELT(cuts, j));
    double* vals = (double*)R_alloc((size_t)off[d] + 1, sizeof(double));
    for (int j = 0; j < d; ++j) {
        const double* c = REAL(VECTOR_ELT(cuts, j));
        for (int k = 0; k < off[j + 1] - off[j]; ++k) vals[off[j] + k] = c[k];
    }
    const int device = 0;
    /* after bartint_init() the rows of X are sharded over every GPU of the group (one call still drives them all) */
    const int x_device = bartint_group_size() > 1 ? BARTINT_DEVICE_GROUP : device;
    bartint_staged *sc = NULL, *sx = NULL;
    int64_t C = 0;
    if (!Rf_isNull(cand)) {
        C = INTEGER(Rf_getAttrib(cand, R_DimSymbol))[0];
        check(bartint_stage_matrix(REAL(cand), C, d, device, &sc));
    }
    if (!Rf_isNull(X)) {
        const int rc = bartint_stage_matrix(REAL(X), INTEGER(Rf_getAttrib(X, R_DimSymbol))[0], d, x_device, &sx);
        if (rc != BARTINT_OK) { bartint_staged_free(sc); check(rc); }
    }
    int n_int = Rf_asInteger(n_draws_integral);
    if (n_int <= 0 || n_int > S) n_int = S;
    SEXP integ = PROTECT(Rf_allocVector(REALSXP, n_int));
    SEXP ims = PROTECT(Rf_allocVector(REALSXP, 2));
    SEXP cv = PROTECT(sc ? Rf_allocVector(REALSXP, C) : R_NilValue);
    SEXP dm = PROTECT(sx ? Rf_allocVector(REALSXP, S) : R_NilValue);
    SEXP dmv = PROTECT(sx ? Rf_allocVector(REALSXP, 2) : R_NilValue);
    const int has_w = !Rf_isNull(weights);
    const int rc = bartint_design_step_dbarts098(strs, REAL(fits), REAL(x), n, d, off, vals, Rf_asReal(ymin), Rf_asReal(ymax),
                                                 m, S, sx, sc, has_w ? REAL(weights) : NULL, has_w ? XLENGTH(weights) : 0,
                                                 Rf_asInteger(measure), n_int, 0, REAL(integ), REAL(ims),
                                                 sc ? REAL(cv) : NULL, sx ? REAL(dm) : NULL, sx ? REAL(dmv) : NULL);
    bartint_staged_free(sc);
    bartint_staged_free(sx);
    if (rc != BARTINT_OK) { UNPROTECT(5); check(rc); }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SET_VECTOR_ELT(out, 0, integ); SET_VECTOR_ELT(out, 1, ims); SET_VECTOR_ELT(out, 2, cv);
    SET_VECTOR_ELT(out, 3, dm); SET_VECTOR_ELT(out, 4, dmv);
    UNPROTECT(6);
    return out;
}
