/* Synthetic BART posterior generator (CPU, bench/test tooling -- not on the evaluation path).
 *
 * Draws S*m regression trees from the BART tree prior that dbarts uses by default
 * (P(split at depth k) = base * (1+k)^-power, base = .95, power = 2) on the
 * integer cut map of a training design, only ever proposing splits that leave at least one training
 * row on each side (so every leaf owns a training row, as in a dbarts fit, and leaf values can be
 * recovered from savedTreeFits).  Leaf values ~ N(0, sigma_mu^2), sigma_mu = 0.5/(k*sqrt(m)).
 * Deterministic for a given seed (xoshiro256** seeded with splitmix64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t s[4];
} rng_t;

static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_unif(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static int rng_int(rng_t* r, int n) { return (int)(rng_unif(r) * n); }
static double rng_norm(rng_t* r) {
    double u1 = rng_unif(r), u2 = rng_unif(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

typedef struct {
    int64_t n_nodes, cap;
    int64_t* tree_off; /* n_trees + 1 */
    int32_t *var, *cut, *left, *right;
    double* mu;
    int64_t n_trees;
} synth_forest;

static int ensure(synth_forest* f, int64_t extra) {
    if (f->n_nodes + extra <= f->cap) return 0;
    int64_t cap = f->cap * 2 + extra + 1024;
    int32_t* v = realloc(f->var, cap * sizeof(int32_t));   if (!v) return -1; f->var = v;
    int32_t* c = realloc(f->cut, cap * sizeof(int32_t));   if (!c) return -1; f->cut = c;
    int32_t* l = realloc(f->left, cap * sizeof(int32_t));  if (!l) return -1; f->left = l;
    int32_t* r = realloc(f->right, cap * sizeof(int32_t)); if (!r) return -1; f->right = r;
    double* m = realloc(f->mu, cap * sizeof(double));      if (!m) return -1; f->mu = m;
    f->cap = cap;
    return 0;
}

/* grow one tree in pre-order; rows[0..nrows) are the training rows in the current node (permuted in place) */
static int grow(synth_forest* f, rng_t* rng, const uint8_t* codes, int64_t n, int d, int64_t* rows, int64_t nrows,
                int depth, int64_t tree_base, double base, double power, double sigma_mu, int max_depth, int* perm) {
    if (ensure(f, 1)) return -1;
    int64_t me = f->n_nodes++;
    f->var[me] = -1; f->cut[me] = 0; f->left[me] = -1; f->right[me] = -1; f->mu[me] = 0.0;
    int split = 0, sv = 0, sc = 0;
    if (depth < max_depth && nrows >= 2 && rng_unif(rng) < base * pow(1.0 + depth, -power)) {
        /* variables in random order; first one on which the node's rows are not constant */
        for (int j = 0; j < d; ++j) perm[j] = j;
        for (int j = d - 1; j > 0; --j) { int k = rng_int(rng, j + 1); int t = perm[j]; perm[j] = perm[k]; perm[k] = t; }
        for (int jj = 0; jj < d && !split; ++jj) {
            int j = perm[jj];
            int lo = 255, hi = 0;
            for (int64_t a = 0; a < nrows; ++a) { int c = codes[(int64_t)j * n + rows[a]]; if (c < lo) lo = c; if (c > hi) hi = c; }
            if (hi > lo) { split = 1; sv = j; sc = lo + rng_int(rng, hi - lo); } /* cut index in [lo, hi-1]: both sides non-empty */
        }
    }
    if (!split) {
        f->mu[me] = sigma_mu * rng_norm(rng);
        return 0;
    }
    /* partition rows: left = code <= cut */
    int64_t nl = 0;
    for (int64_t a = 0; a < nrows; ++a)
        if (codes[(int64_t)sv * n + rows[a]] <= sc) { int64_t t = rows[a]; rows[a] = rows[nl]; rows[nl] = t; ++nl; }
    f->var[me] = sv; f->cut[me] = sc;
    f->left[me] = (int32_t)(f->n_nodes - tree_base);
    if (grow(f, rng, codes, n, d, rows, nl, depth + 1, tree_base, base, power, sigma_mu, max_depth, perm)) return -1;
    f->right[me] = (int32_t)(f->n_nodes - tree_base);
    if (grow(f, rng, codes, n, d, rows + nl, nrows - nl, depth + 1, tree_base, base, power, sigma_mu, max_depth, perm)) return -1;
    return 0;
}

/* train_codes: feature-major uint8 [j * n + i] (cut-map of the training design) */
synth_forest* bartint_synth_create(uint64_t seed, int32_t S, int32_t m, int32_t d, int64_t n, const uint8_t* train_codes,
                                   double base, double power, double sigma_mu, int32_t max_depth) {
    synth_forest* f = calloc(1, sizeof(synth_forest));
    if (!f) return NULL;
    f->n_trees = (int64_t)S * m;
    f->tree_off = calloc((size_t)f->n_trees + 1, sizeof(int64_t));
    int64_t* rows = malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    int* perm = malloc((size_t)(d > 0 ? d : 1) * sizeof(int));
    if (!f->tree_off || !rows || !perm) goto fail;
    rng_t rng;
    rng_seed(&rng, seed);
    for (int64_t k = 0; k < f->n_trees; ++k) {
        for (int64_t i = 0; i < n; ++i) rows[i] = i;
        f->tree_off[k] = f->n_nodes;
        if (grow(f, &rng, train_codes, n, d, rows, n, 0, f->n_nodes, base, power, sigma_mu, max_depth, perm)) goto fail;
    }
    f->tree_off[f->n_trees] = f->n_nodes;
    free(rows); free(perm);
    return f;
fail:
    free(rows); free(perm);
    if (f) { free(f->tree_off); free(f->var); free(f->cut); free(f->left); free(f->right); free(f->mu); free(f); }
    return NULL;
}

int64_t bartint_synth_nodes(const synth_forest* f) { return f ? f->n_nodes : -1; }

void bartint_synth_copy(const synth_forest* f, int64_t* tree_off, int32_t* var, int32_t* cut, int32_t* left,
                        int32_t* right, double* mu) {
    memcpy(tree_off, f->tree_off, (size_t)(f->n_trees + 1) * sizeof(int64_t));
    memcpy(var, f->var, (size_t)f->n_nodes * sizeof(int32_t));
    memcpy(cut, f->cut, (size_t)f->n_nodes * sizeof(int32_t));
    memcpy(left, f->left, (size_t)f->n_nodes * sizeof(int32_t));
    memcpy(right, f->right, (size_t)f->n_nodes * sizeof(int32_t));
    memcpy(mu, f->mu, (size_t)f->n_nodes * sizeof(double));
}

void bartint_synth_free(synth_forest* f) {
    if (!f) return;
    free(f->tree_off); free(f->var); free(f->cut); free(f->left); free(f->right); free(f->mu); free(f);
}
