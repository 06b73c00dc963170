/* CPU ORACLE, C restatement (test infrastructure + the bench's CPU baseline; NOT product code).
 *
 * PARITY UNPINNED against real dbarts output: the reference ships no golden vectors for this path and R
 * is not installable here (see oracle/oracle_np.py header).  This file restates the same arithmetic as
 * oracle_np.py in plain C so that (a) larger parity cases finish in seconds and (b) bench.py can time
 * "the reference's CPU implementation of the path" on the GPU box's host cores:
 *
 *   oracle_predict_reduce   dbarts predict loop [dbarts-mem] (for each saved draw, for each tree, for all
 *                           rows: walk from the root, right iff x[var] > cut value, add leaf mu; then
 *                           (ymax-ymin)*(sum+0.5)+ymin) with rowMeans (bartMean.R:76) and colVars
 *                           (BARTInt.R:314) folded in so the S x N matrix need not be stored.
 *                           OpenMP over draws = the doParallel-style split BASELINE.md describes; dbarts
 *                           itself runs single-threaded in the reference (nthread=1 default).
 *   oracle_exact_integrals  sampleIntegrals chain, BARTInt.R:10-223 (all three measures, as written).
 *   oracle_leaf_ordinals    leaf numbering left-to-right in pre-order.
 *
 * Forest layout = the flat SoA of oracle_np.py (child indices relative to the tree's first node).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t S, m, d;
    const int64_t* tree_offset;
    const int32_t *var, *cut, *left, *right;
    const double* mu;
    const int32_t* cut_offsets;
    const double* cut_values;
    double y_min, y_max;
    int32_t route;   /* 0: dbarts, right iff x > cut (NaN -> left); 1: BART::wbart tree::bn, left iff x < cut (NaN -> right) */
} forest_t;

static inline int go_right(int32_t route, double x, double c) { return route ? !(x < c) : (x > c); }

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* one tree over all rows, dbarts order: tree outer, rows inner; acc[i] += mu(leaf(x_i)) */
static void tree_predict_add(const forest_t* f, int64_t k, const double* X, int64_t N, double* acc) {
    const int64_t a = f->tree_offset[k];
    const int32_t *var = f->var + a, *cut = f->cut + a, *left = f->left + a, *right = f->right + a;
    const double* mu = f->mu + a;
    if (var[0] < 0) {
        for (int64_t i = 0; i < N; ++i) acc[i] += mu[0];
        return;
    }
    for (int64_t i = 0; i < N; ++i) {
        int32_t node = 0;
        while (var[node] >= 0) {
            const double c = f->cut_values[f->cut_offsets[var[node]] + cut[node]];
            node = go_right(f->route, X[(int64_t)var[node] * N + i], c) ? right[node] : left[node];
        }
        acc[i] += mu[node];
    }
}

/* X: N x d column-major.  Outputs (any may be NULL): draw_mean[S] = rowMeans(predict), point_mean[N] =
 * colMeans, point_var[N] = colVars (n-1), full[S x N col-major, rows = draws] = predict(model, X). */
void oracle_predict_reduce(int32_t S, int32_t m, int32_t d, const int64_t* tree_offset, const int32_t* var,
                           const int32_t* cut, const int32_t* left, const int32_t* right, const double* mu,
                           const int32_t* cut_offsets, const double* cut_values, double y_min, double y_max,
                           int32_t route, const double* X, int64_t N, int32_t nthreads, double* draw_mean,
                           double* point_mean, double* point_var, double* full) {
    forest_t f = {S, m, d, tree_offset, var, cut, left, right, mu, cut_offsets, cut_values, y_min, y_max, route};
    const int want_pt = point_mean != NULL || point_var != NULL;
    if (nthreads < 1) nthreads = 1;
    /* per-thread Welford state over the draws it owns, merged afterwards in thread order */
    double* cnt = calloc((size_t)nthreads, sizeof(double));
    double* pmean = want_pt ? calloc((size_t)nthreads * (size_t)(N > 0 ? N : 1), sizeof(double)) : NULL;
    double* pm2 = want_pt ? calloc((size_t)nthreads * (size_t)(N > 0 ? N : 1), sizeof(double)) : NULL;
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* acc = malloc((size_t)(N > 0 ? N : 1) * sizeof(double));
        double* mymean = want_pt ? pmean + (size_t)tid * (size_t)N : NULL;
        double* mym2 = want_pt ? pm2 + (size_t)tid * (size_t)N : NULL;
#pragma omp for schedule(static)
        for (int32_t s = 0; s < S; ++s) {
            for (int64_t i = 0; i < N; ++i) acc[i] = 0.0;
            for (int32_t t = 0; t < m; ++t) tree_predict_add(&f, (int64_t)s * m + t, X, N, acc);
            long double rs = 0.0L;
            cnt[tid] += 1.0;
            const double c = cnt[tid];
            for (int64_t i = 0; i < N; ++i) {
                const double v = (y_max - y_min) * (acc[i] + 0.5) + y_min;
                rs += v;
                if (full) full[(int64_t)s + (int64_t)S * i] = v;
                if (want_pt) {
                    const double dlt = v - mymean[i];
                    mymean[i] += dlt / c;
                    mym2[i] += dlt * (v - mymean[i]);
                }
            }
            if (draw_mean) draw_mean[s] = (double)(rs / (long double)N);
        }
        free(acc);
    }
    if (want_pt) {
        for (int64_t i = 0; i < N; ++i) {
            double n = 0.0, mean = 0.0, m2 = 0.0;
            for (int t = 0; t < nthreads; ++t) {
                const double nb = cnt[t];
                if (nb == 0.0) continue;
                const double mb = pmean[(size_t)t * (size_t)N + i], m2b = pm2[(size_t)t * (size_t)N + i];
                if (n == 0.0) { n = nb; mean = mb; m2 = m2b; continue; }
                const double tot = n + nb, delta = mb - mean;
                mean += delta * (nb / tot);
                m2 += m2b + delta * delta * (n * nb / tot);
                n = tot;
            }
            if (point_mean) point_mean[i] = mean;
            if (point_var) point_var[i] = m2 / (n - 1.0);
        }
    }
    free(cnt); free(pmean); free(pm2);
}

void oracle_leaf_ordinals(int32_t S, int32_t m, const int64_t* tree_offset, const int32_t* var, const int32_t* cut,
                          const int32_t* left, const int32_t* right, const int32_t* cut_offsets,
                          const double* cut_values, int32_t route, const double* X, int64_t N, int32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)S * m; ++k) {
        const int64_t a = tree_offset[k];
        const int32_t cnt = (int32_t)(tree_offset[k + 1] - a);
        int32_t* ordinal = malloc((size_t)cnt * sizeof(int32_t));
        int32_t* stack = malloc((size_t)cnt * sizeof(int32_t));
        int sp = 0, ord = 0;
        stack[sp++] = 0;
        while (sp > 0) {            /* pre-order visit, leaves numbered in visiting order */
            int32_t n = stack[--sp];
            if (var[a + n] < 0) ordinal[n] = ord++;
            else { stack[sp++] = right[a + n]; stack[sp++] = left[a + n]; }
        }
        for (int64_t i = 0; i < N; ++i) {
            int32_t node = 0;
            while (var[a + node] >= 0) {
                const double c = cut_values[cut_offsets[var[a + node]] + cut[a + node]];
                node = go_right(route, X[(int64_t)var[a + node] * N + i], c) ? right[a + node] : left[a + node];
            }
            out[k * N + i] = ordinal[node];
        }
        free(ordinal); free(stack);
    }
}

/* ---- exact integral: BARTInt.R:32-72 (fillProbabilityForNode), :10-30 (terminalProbability), :165-173 ---- */
static double pnorm_(double x) { return 0.5 * erfc(-x / sqrt(2.0)); }
static double pexp_(double q) { return q <= 0.0 ? 0.0 : (isinf(q) ? 1.0 : -expm1(-q)); }

typedef struct {
    const int32_t *var, *cut, *left, *right;
    const double* mu;
    const forest_t* f;
    int measure;
    double* prob;      /* conditional probability per node */
    int32_t* parent;
    long double total;
} tree_ctx;

static void fill_prob(tree_ctx* c, int32_t k, double* box /* [2*d]: lo then hi */) {
    const int d = c->f->d;
    if (c->left[k] >= 0) {
        const int32_t v = c->var[k];
        const double rule = c->f->cut_values[c->f->cut_offsets[v] + c->cut[k]];
        const double lo = box[v], hi = box[d + v];
        double pl, pr;
        if (c->measure == 0) { pl = (rule - lo) / (hi - lo); pr = (hi - rule) / (hi - lo); }
        else if (c->measure == 1) {
            const double norm = pnorm_(hi - 0.5) - pnorm_(lo - 0.5);
            pl = (pnorm_(rule) - pnorm_(lo)) / norm; pr = 1.0 - pl;
        } else if (c->measure == 3) {   /* corrected gaussian (not the reference): both masses about mean 0.5 */
            const double norm = pnorm_(hi - 0.5) - pnorm_(lo - 0.5);
            pl = (pnorm_(rule - 0.5) - pnorm_(lo - 0.5)) / norm; pr = (pnorm_(hi - 0.5) - pnorm_(rule - 0.5)) / norm;
        } else {
            const double norm = pexp_(hi) - pexp_(lo);
            pl = (pexp_(rule) - pexp_(lo)) / norm; pr = 1.0 - pl;
        }
        c->prob[c->left[k]] = pl; c->prob[c->right[k]] = pr;
        c->parent[c->left[k]] = k; c->parent[c->right[k]] = k;
        box[d + v] = rule; fill_prob(c, c->left[k], box); box[d + v] = hi;
        box[v] = rule;     fill_prob(c, c->right[k], box); box[v] = lo;
    } else if (k == 0) {
        c->prob[0] = 1.0;   /* stump */
    }
}

static void sum_leaves(tree_ctx* c, int32_t k, double* acc) {
    if (c->left[k] >= 0) { sum_leaves(c, c->left[k], acc); sum_leaves(c, c->right[k], acc); return; }
    double p = c->prob[k];
    int32_t node = k;
    while (c->parent[node] > 0) { node = c->parent[node]; p = p * c->prob[node]; }
    *acc = *acc + p * c->mu[k];
}

void oracle_exact_integrals(int32_t S, int32_t m, int32_t d, const int64_t* tree_offset, const int32_t* var,
                            const int32_t* cut, const int32_t* left, const int32_t* right, const double* mu,
                            const int32_t* cut_offsets, const double* cut_values, int32_t measure, const double* lo,
                            const double* hi, int32_t n_draws, int32_t nthreads, double* out) {
    forest_t f = {S, m, d, tree_offset, var, cut, left, right, mu, cut_offsets, cut_values, 0, 1};
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int32_t s = 0; s < n_draws; ++s) {
        long double acc = 0.0L;          /* posteriorSum uses base::sum (long double) */
        double* box = malloc((size_t)(2 * (d > 0 ? d : 1)) * sizeof(double));
        for (int32_t t = 0; t < m; ++t) {
            const int64_t k = (int64_t)s * m + t, a = tree_offset[k];
            const int32_t cnt = (int32_t)(tree_offset[k + 1] - a);
            for (int j = 0; j < d; ++j) {
                box[j] = lo ? lo[j] : 0.0;
                box[d + j] = hi ? hi[j] : (measure == 2 ? INFINITY : 1.0);
            }
            tree_ctx c = {var + a, cut + a, left + a, right + a, mu + a, &f, measure,
                          malloc((size_t)cnt * sizeof(double)), malloc((size_t)cnt * sizeof(int32_t)), 0.0L};
            for (int32_t q = 0; q < cnt; ++q) c.parent[q] = -1;
            fill_prob(&c, 0, box);
            double tree_sum = 0.0;
            sum_leaves(&c, 0, &tree_sum);
            acc += tree_sum;
            free(c.prob); free(c.parent);
        }
        out[s] = (double)acc;
        free(box);
    }
}
