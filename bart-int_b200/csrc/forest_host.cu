// Host side of the forest container: the dbarts 0.9-8 tree-state exporter (strings + savedTreeFits ->
// SoA), SoA canonicalisation, table-group planning and the upload to HBM.
//
// Reference behaviour restated here (never copied): getTree(), src/BARTInt.R:92-133 --
//   :110-111 savedTrees[treeNum, sampleNum] / savedTreeFits[, treeNum, sampleNum]
//   :118-119 tokenisation gsub("\\.", "\\. ") + strsplit(" ") feeding dbarts:::buildTree
//   :122-125 route the training rows, leaf mu = fit of a training row in that leaf
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <ctime>

#include "common.cuh"

namespace bartint {

static thread_local char g_err[512] = "";

// BARTINT_TIMING=1: print the phases of forest construction to stderr (host wall clock)
struct PhaseTimer {
    bool on;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    PhaseTimer() : on(getenv("BARTINT_TIMING") != nullptr), t0(now()) {}
    void lap(const char* what) {                       // "" restarts the clock without printing
        if (!on) return;
        const double t = now();
        if (what[0]) fprintf(stderr, "[bartint] %-28s %8.3f ms\n", what, t - t0);
        t0 = t;
    }
};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    // clear the sticky-less error state so later calls report their own failures
    (void)cudaGetLastError();
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? BARTINT_ERR_NODEVICE : BARTINT_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------
// string parser: tree := "." | VAR " " CUT " " tree tree   (pre-order, 0-based ints)
// Accepts any run of spaces between tokens (the reference tokeniser inserts one after every '.').
// Appends the tree's nodes in pre-order; child indices are relative to the tree's first node.
// ---------------------------------------------------------------------------------------------
static int forest_check_cuts(bartint_forest* f);      // defined with forest_finish_host

static bool parse_int(const char*& p, int32_t& out) {
    while (*p == ' ') ++p;
    if (*p < '0' || *p > '9') return false;
    int64_t v = 0;
    while (*p >= '0' && *p <= '9') {
        v = v * 10 + (*p - '0');
        if (v > std::numeric_limits<int32_t>::max()) return false;
        ++p;
    }
    out = (int32_t)v;
    return true;
}

// Parses one tree string straight into the SoA slices of its tree (capacity `cap` nodes, known from the count of
// '.' characters).  `pending` is a scratch stack of internal nodes still waiting for their right child.
static int parse_tree_string(const char* s, int32_t cap, int32_t* var, int32_t* cut, int32_t* left, int32_t* right,
                             std::vector<int32_t>& pending) {
    const char* p = s;
    pending.clear();
    int32_t n = 0;
    bool done = false;
    while (!done) {
        while (*p == ' ') ++p;
        if (*p == '\0' || n >= cap) return BARTINT_ERR_PARSE;
        const int32_t me = n++;
        if (*p == '.') {
            ++p;
            var[me] = -1; cut[me] = 0; left[me] = -1; right[me] = -1;
            // a finished subtree: the left child of the node above (already linked as parent+1), or the right child of
            // the innermost pending node
            while (true) {
                if (pending.empty()) { done = true; break; }
                const int32_t par = pending.back();
                if (right[par] == -2) {            // left subtree just completed -> the next node is its right child
                    right[par] = n;
                    break;
                }
                pending.pop_back();                // right subtree completed -> this node is complete too
            }
        } else {
            int32_t v, c;
            if (!parse_int(p, v) || !parse_int(p, c)) return BARTINT_ERR_PARSE;
            var[me] = v; cut[me] = c; left[me] = me + 1;
            right[me] = -2;                        // -2: waiting for the left subtree to finish
            pending.push_back(me);
        }
    }
    while (*p == ' ') ++p;
    if (*p != '\0' || n != cap) return BARTINT_ERR_PARSE;     // tree$remainder must be empty
    return BARTINT_OK;
}

static inline int bin_value(double x, const double* cuts, int nc) {
    // #{k : x > cut[k]} for ascending cuts; NaN -> 0
    int lo = 0, len = nc;
    while (len > 0) {
        int h = len >> 1;
        bool r = cuts[lo + h] < x;
        lo = r ? lo + h + 1 : lo;
        len = r ? len - h - 1 : h;
    }
    return lo;
}

int parse_dbarts_forest(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                        int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                        int32_t m, int32_t S, bartint_forest* f) {
    const int64_t n_trees = (int64_t)S * m;
    PhaseTimer pt;
    // pass 1: a tree with K leaves ('.' characters) has 2K-1 nodes -> node offsets by prefix sum
    f->tree_off.assign((size_t)n_trees + 1, 0);
    int64_t bad_tree = -1;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_trees; ++k) {
        int64_t dots = 0;
        if (tree_strings[k] != nullptr)
            for (const char* p = tree_strings[k]; *p; ++p) dots += (*p == '.');
        if (dots == 0) {
#pragma omp critical
            bad_tree = k;
        }
        f->tree_off[(size_t)k + 1] = 2 * dots - 1;
    }
    if (bad_tree < 0)
        for (int64_t k = 0; k < n_trees; ++k) f->tree_off[(size_t)k + 1] += f->tree_off[(size_t)k];
    // integer cut map of the training matrix (dbarts keeps x as cut indices internally)
    std::vector<int32_t> codes((size_t)(n * d));
#pragma omp parallel for schedule(static)
    for (int32_t j = 0; j < d; ++j) {
        const double* cuts = cut_values + cut_offsets[j];
        const int nc = cut_offsets[j + 1] - cut_offsets[j];
        for (int64_t i = 0; i < n; ++i) codes[(size_t)(j * n + i)] = bin_value(x_train[j * n + i], cuts, nc);
    }
    // rows of the training matrix ordered by code, per variable (candidate rows for the leaf values below)
    constexpr int64_t kLeafTries = 8;
    std::vector<int32_t> order((size_t)(n * d));
#pragma omp parallel for schedule(static)
    for (int32_t j = 0; j < d; ++j) {
        int32_t* o = order.data() + (size_t)j * (size_t)n;
        const int32_t* cj = codes.data() + (size_t)j * (size_t)n;
        const int nc = cut_offsets[j + 1] - cut_offsets[j];          // codes are 0 .. nc: stable counting sort
        std::vector<int64_t> first((size_t)nc + 2, 0);
        for (int64_t i = 0; i < n; ++i) ++first[(size_t)cj[i] + 1];
        for (int c = 0; c <= nc; ++c) first[(size_t)c + 1] += first[(size_t)c];
        for (int64_t i = 0; i < n; ++i) o[first[(size_t)cj[i]]++] = (int32_t)i;
    }
    const int64_t nn = bad_tree < 0 ? f->tree_off[(size_t)n_trees] : 0;
    f->var.resize((size_t)nn); f->cut.resize((size_t)nn); f->left.resize((size_t)nn); f->right.resize((size_t)nn);
    f->mu.resize((size_t)nn);
    int bad_ref = 0;
    pt.lap("exporter: count + cut map");
    // pass 2: parse in place, then recover the leaf values
#pragma omp parallel
    {
        std::vector<int32_t> pending;
        std::vector<char> have;
        std::vector<std::pair<double*, const double*>> pend;     // leaf value reads in flight: (destination, source)
        auto drain = [&pend](size_t keep) {                      // collect all but the `keep` youngest
            const size_t nd = pend.size() > keep ? pend.size() - keep : 0;
            for (size_t q = 0; q < nd; ++q) *pend[q].first = *pend[q].second;
            pend.erase(pend.begin(), pend.begin() + (std::ptrdiff_t)nd);
        };
#pragma omp for schedule(static) nowait
        for (int64_t k = 0; k < n_trees; ++k) {
            if (bad_tree >= 0) continue;
            const int64_t base = f->tree_off[(size_t)k];
            const int32_t cnt = (int32_t)(f->tree_off[(size_t)k + 1] - base);
            int32_t* var = f->var.data() + base;
            int32_t* cut = f->cut.data() + base;
            int32_t* left = f->left.data() + base;
            int32_t* right = f->right.data() + base;
            double* mu = f->mu.data() + base;
            if (parse_tree_string(tree_strings[k], cnt, var, cut, left, right, pending) != BARTINT_OK) {
#pragma omp critical
                bad_tree = k;
                continue;
            }
            bool ok = true;
            for (int32_t a = 0; a < cnt; ++a) {
                mu[a] = var[a] < 0 ? std::numeric_limits<double>::quiet_NaN() : 0.0;
                if (var[a] >= 0 && (var[a] >= d || cut[a] >= cut_offsets[var[a] + 1] - cut_offsets[var[a]])) ok = false;
            }
            if (!ok) {
#pragma omp atomic write
                bad_ref = 1;
                continue;
            }
            // leaf mu = fit of the first training row routed to the leaf (fillPlotInfoForNode, BARTInt.R:125)
            const double* fits = tree_fits + (size_t)k * (size_t)n;   // [obs, tree, sample] column-major
            const int32_t n_leaf = (cnt + 1) / 2;
            // Which row?  Any row of the leaf carries the leaf's value.  A leaf that is the left (right) child of a
            // split on variable v holds rows with code[v] <= cut (> cut), so the rows with the smallest (largest)
            // codes of v are tried first and verified by routing them from the root: one or two short walks and one
            // read of `fits` per leaf, instead of routing the rows in order until every leaf has turned up (which
            // also dragged most of the n x m x S fits array through the cache).
            have.assign((size_t)cnt, 0);
            int32_t found = 0;
            if (pend.size() >= 192) drain(64);
            if (cnt == 1) {
                if (n > 0) { __builtin_prefetch(fits); pend.emplace_back(mu, fits); found = 1; }
            } else {
                for (int32_t p = 0; p < cnt; ++p) {
                    if (var[p] < 0) continue;
                    const int32_t v = var[p], c = cut[p];
                    const int32_t* ord = order.data() + (size_t)v * (size_t)n;
                    const int32_t* cv = codes.data() + (size_t)v * (size_t)n;
                    for (int side = 0; side < 2; ++side) {
                        const int32_t leaf = side ? right[p] : left[p];
                        if (var[leaf] >= 0) continue;
                        for (int64_t t = 0; t < n && t < kLeafTries; ++t) {
                            const int64_t i = side ? ord[n - 1 - t] : ord[t];
                            if ((cv[i] > c) != (side == 1)) break;          // no row of this leaf's side is left
                            int32_t a = 0;
                            while (var[a] >= 0) a = (codes[(size_t)(var[a] * n + i)] > cut[a]) ? right[a] : left[a];
                            if (a == leaf) {
                                // the read of fits[i] is a cache miss (the array is n x m x S doubles, touched once
                                // per leaf): start it now, collect it later, many misses in flight at once
                                __builtin_prefetch(fits + i);
                                pend.emplace_back(mu + a, fits + i);
                                have[(size_t)a] = 1; ++found;
                                break;
                            }
                        }
                    }
                }
            }
            // leaves the candidates missed (boxes that are narrow in an earlier split): rows in order
            for (int64_t i = 0; i < n && found < n_leaf; ++i) {
                int32_t a = 0;
                while (var[a] >= 0) a = (codes[(size_t)(var[a] * n + i)] > cut[a]) ? right[a] : left[a];
                if (!have[(size_t)a]) { mu[a] = fits[i]; have[(size_t)a] = 1; ++found; }
            }
        }
        drain(0);
    }
    pt.lap("exporter: parse + leaf mu");
    if (bad_tree >= 0) {
        set_error("malformed dbarts tree string at tree %lld (draw %lld)", (long long)(bad_tree % m),
                  (long long)(bad_tree / m));
        return BARTINT_ERR_PARSE;
    }
    if (bad_ref) {
        set_error("tree string references a variable or cut index outside the cut-point table");
        return BARTINT_ERR_PARSE;
    }
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// The exporter of the fused design step: dbarts 0.9-8 strings + fits -> WALK FORM in one pass.  The trees of a design
// step are evaluated (bit-sliced sums, node-walk moments, exact integrals) and thrown away, so the canonical SoA that
// parse_dbarts_forest fills, forest_finish_host re-reads for its statistics and pack_walk_form re-reads to pack never
// needs to exist: every tree is parsed into a small thread-local scratch, validated, measured (depth classes, leaves)
// and written straight into the page-locked walk form; the leaf values are collected through the same prefetch queue.
// Same grammar, same checks and the same leaf-value rule (getTree, src/BARTInt.R:92-133) as parse_dbarts_forest; the
// GPU parity tests of the design step compare the two paths.
// ---------------------------------------------------------------------------------------------
int parse_dbarts_walk(const char* const* tree_strings, const double* tree_fits, const double* x_train, int64_t n, int32_t d,
                      const int32_t* cut_offsets, const double* cut_values, int32_t m, int32_t S, bartint_forest* f,
                      WalkForm& W) {
    const int64_t n_trees = (int64_t)S * m;
    PhaseTimer pt;
    int rc = forest_check_cuts(f);
    if (rc) return rc;
    W.tree_off.resize((size_t)n_trees + 1);
    W.leaf_off.resize((size_t)n_trees + 1);
    int64_t bad_tree = -1;
    std::vector<int32_t> n_leaf((size_t)n_trees);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_trees; ++k) {
        int32_t dots = 0;
        if (tree_strings[k] != nullptr)
            for (const char* p = tree_strings[k]; *p; ++p) dots += (*p == '.');
        if (dots == 0) {
#pragma omp critical
            bad_tree = k;
        }
        n_leaf[(size_t)k] = dots;
    }
    if (bad_tree >= 0) {
        set_error("malformed dbarts tree string at tree %lld (draw %lld)", (long long)(bad_tree % m), (long long)(bad_tree / m));
        return BARTINT_ERR_PARSE;
    }
    int64_t nn = 0, nl = 0;
    int32_t draw_internal_max = 0;
    for (int32_t s = 0; s < S; ++s) {
        int64_t internal = 0;
        for (int32_t t = 0; t < m; ++t) {
            const size_t k = (size_t)s * (size_t)m + (size_t)t;
            W.tree_off[k] = (int32_t)nn; W.leaf_off[k] = (int32_t)nl;
            nn += 2 * (int64_t)n_leaf[k] - 1; nl += n_leaf[k];
            internal += n_leaf[k] - 1;
        }
        draw_internal_max = std::max<int64_t>(draw_internal_max, internal);
    }
    BI_REQUIRE(nn < (int64_t)1 << 31, BARTINT_ERR_UNSUPPORTED, "forest has more than 2^31 nodes");
    W.tree_off[(size_t)n_trees] = (int32_t)nn; W.leaf_off[(size_t)n_trees] = (int32_t)nl;
    W.nodes.resize((size_t)nn);
    W.leaf_mu.resize((size_t)nl);
    // integer cut map of the training matrix and its rows ordered by code (see parse_dbarts_forest)
    std::vector<int32_t> codes((size_t)(n * d)), order((size_t)(n * d));
#pragma omp parallel for schedule(static)
    for (int32_t j = 0; j < d; ++j) {
        const double* cuts = cut_values + cut_offsets[j];
        const int nc = cut_offsets[j + 1] - cut_offsets[j];
        int32_t* cj = codes.data() + (size_t)j * (size_t)n;
        for (int64_t i = 0; i < n; ++i) cj[i] = bin_value(x_train[j * n + i], cuts, nc);
        int32_t* o = order.data() + (size_t)j * (size_t)n;
        std::vector<int64_t> first((size_t)nc + 2, 0);
        for (int64_t i = 0; i < n; ++i) ++first[(size_t)cj[i] + 1];
        for (int c = 0; c <= nc; ++c) first[(size_t)c + 1] += first[(size_t)c];
        for (int64_t i = 0; i < n; ++i) o[first[(size_t)cj[i]]++] = (int32_t)i;
    }
    pt.lap("exporter: count + cut map");
    constexpr int64_t kLeafTries = 8;
    int bad_ref = 0;
    int32_t max_depth = 0, max_internal = 0;
    int64_t cls0 = 0, cls1 = 0, cls2 = 0, cls3 = 0, cls4 = 0, refs = 0;
#pragma omp parallel reduction(max : max_depth, max_internal) reduction(+ : cls0, cls1, cls2, cls3, cls4, refs)
    {
        std::vector<int32_t> pending, var, cut, left, right, depth, ord_of;
        std::vector<char> have;
        std::vector<std::pair<double*, const double*>> pend;     // leaf value reads in flight: (destination, source)
        auto drain = [&pend](size_t keep) {
            const size_t nd = pend.size() > keep ? pend.size() - keep : 0;
            for (size_t q = 0; q < nd; ++q) *pend[q].first = *pend[q].second;
            pend.erase(pend.begin(), pend.begin() + (std::ptrdiff_t)nd);
        };
#pragma omp for schedule(static) nowait
        for (int64_t k = 0; k < n_trees; ++k) {
            if (bad_tree >= 0) continue;
            const int32_t cnt = 2 * n_leaf[(size_t)k] - 1;
            if ((int32_t)var.size() < cnt) {
                var.resize((size_t)cnt); cut.resize((size_t)cnt); left.resize((size_t)cnt); right.resize((size_t)cnt);
                depth.resize((size_t)cnt); ord_of.resize((size_t)cnt);
            }
            if (parse_tree_string(tree_strings[k], cnt, var.data(), cut.data(), left.data(), right.data(), pending) != BARTINT_OK) {
#pragma omp critical
                bad_tree = k;
                continue;
            }
            uint2* nodes = W.nodes.data() + W.tree_off[(size_t)k];
            double* mu = W.leaf_mu.data() + W.leaf_off[(size_t)k];
            bool ok = true;
            int32_t ord = 0, internal = 0;
            depth[0] = 0;
            for (int32_t a = 0; a < cnt; ++a) {
                if (var[(size_t)a] >= 0) {
                    const int32_t v = var[(size_t)a], c = cut[(size_t)a];
                    if (v >= d || c >= cut_offsets[v + 1] - cut_offsets[v]) { ok = false; break; }
                    nodes[a] = make_uint2((uint32_t)v | ((uint32_t)c << 16), (uint32_t)(right[(size_t)a] - a));
                    const int32_t dp = depth[(size_t)a];
                    depth[(size_t)left[(size_t)a]] = dp + 1; depth[(size_t)right[(size_t)a]] = dp + 1;
                    ++internal;
                    if (dp == 0) ++cls0; else if (dp == 1) ++cls1; else if (dp == 2) ++cls2; else if (dp == 3) ++cls3; else ++cls4;
                    if (dp >= 1) refs += dp + 1;
                } else {
                    nodes[a] = make_uint2(LEAF_VAR, (uint32_t)ord);
                    ord_of[(size_t)a] = ord;
                    mu[ord] = std::numeric_limits<double>::quiet_NaN();
                    ++ord;
                }
                max_depth = std::max(max_depth, depth[(size_t)a]);
            }
            if (!ok) {
#pragma omp atomic write
                bad_ref = 1;
                continue;
            }
            max_internal = std::max(max_internal, internal);
            // leaf mu = fit of a training row routed to the leaf (fillPlotInfoForNode, BARTInt.R:125); candidate rows
            // from the per-variable code order, verified by routing them from the root (see parse_dbarts_forest)
            const double* fits = tree_fits + (size_t)k * (size_t)n;
            const int32_t nlf = (cnt + 1) / 2;
            have.assign((size_t)cnt, 0);
            int32_t found = 0;
            if (pend.size() >= 192) drain(64);
            if (cnt == 1) {
                if (n > 0) { __builtin_prefetch(fits); pend.emplace_back(mu, fits); found = 1; }
            } else {
                for (int32_t p = 0; p < cnt; ++p) {
                    if (var[(size_t)p] < 0) continue;
                    const int32_t v = var[(size_t)p], c = cut[(size_t)p];
                    const int32_t* o = order.data() + (size_t)v * (size_t)n;
                    const int32_t* cv = codes.data() + (size_t)v * (size_t)n;
                    for (int side = 0; side < 2; ++side) {
                        const int32_t leaf = side ? right[(size_t)p] : left[(size_t)p];
                        if (var[(size_t)leaf] >= 0) continue;
                        for (int64_t t = 0; t < n && t < kLeafTries; ++t) {
                            const int64_t i = side ? o[n - 1 - t] : o[t];
                            if ((cv[i] > c) != (side == 1)) break;
                            int32_t a = 0;
                            while (var[(size_t)a] >= 0)
                                a = (codes[(size_t)(var[(size_t)a] * n + i)] > cut[(size_t)a]) ? right[(size_t)a] : left[(size_t)a];
                            if (a == leaf) {
                                __builtin_prefetch(fits + i);
                                pend.emplace_back(mu + ord_of[(size_t)a], fits + i);
                                have[(size_t)a] = 1; ++found;
                                break;
                            }
                        }
                    }
                }
            }
            for (int64_t i = 0; i < n && found < nlf; ++i) {        // leaves the candidates missed: rows in order
                int32_t a = 0;
                while (var[(size_t)a] >= 0)
                    a = (codes[(size_t)(var[(size_t)a] * n + i)] > cut[(size_t)a]) ? right[(size_t)a] : left[(size_t)a];
                if (!have[(size_t)a]) { mu[ord_of[(size_t)a]] = fits[i]; have[(size_t)a] = 1; ++found; }
            }
        }
        drain(0);
    }
    if (bad_tree >= 0) {
        set_error("malformed dbarts tree string at tree %lld (draw %lld)", (long long)(bad_tree % m), (long long)(bad_tree / m));
        return BARTINT_ERR_PARSE;
    }
    if (bad_ref) {
        set_error("tree string references a variable or cut index outside the cut-point table");
        return BARTINT_ERR_PARSE;
    }
    double max_abs_mu = 0.0;
    bool mu_finite = true;
#pragma omp parallel for schedule(static) reduction(max : max_abs_mu) reduction(&& : mu_finite)
    for (int64_t g = 0; g < nl; ++g) {
        const double a = std::fabs(W.leaf_mu[(size_t)g]);
        if (!(a <= 1.7e308)) mu_finite = false;
        else max_abs_mu = std::max(max_abs_mu, a);
    }
    f->n_nodes = nn; f->n_leaves = nl; f->max_depth = max_depth; f->max_internal = max_internal;
    f->bs_class_nodes[0] = cls0; f->bs_class_nodes[1] = cls1; f->bs_class_nodes[2] = cls2; f->bs_class_nodes[3] = cls3;
    f->bs_class_nodes[4] = cls4; f->bs_refs_total = refs;
    f->draw_internal_max = draw_internal_max;
    f->max_abs_mu = mu_finite ? max_abs_mu : INFINITY;
    f->walk_only = true;
    bs_configure(f);
    pt.lap("exporter: parse + leaf mu + walk form");
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// BART::wbart / pbart `treedraws$trees` text:
//   "ndpost ntree p\n"  then per draw, per tree:  "nnodes\n"  followed by nnodes lines  "nid var cut theta\n"
// (tree.cpp operator<<: nodes as tree::getnodes lists them).  nid = heap index (root 1, children 2 nid, 2 nid + 1);
// a node is internal iff its left child is listed; var / cut 0-based; theta = leaf value.
// Pass 1 (sequential) finds every tree's node lines by counting newlines; pass 2 (OpenMP over trees) parses the
// lines and emits the canonical pre-order SoA.
// ---------------------------------------------------------------------------------------------
static bool parse_i64(const char*& p, const char* end, int64_t& out) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
    if (p >= end || *p < '0' || *p > '9') return false;
    int64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        if (v > (std::numeric_limits<int64_t>::max() - 9) / 10) return false;
        v = v * 10 + (*p - '0');
        ++p;
    }
    out = v;
    return true;
}

int parse_wbart_forest(const char* text, int64_t len, int32_t d, bartint_forest* f) {
    const char* p = text;
    const char* end = text + len;
    int64_t S64 = 0, m64 = 0, p64 = 0;
    BI_REQUIRE(parse_i64(p, end, S64) && parse_i64(p, end, m64) && parse_i64(p, end, p64), BARTINT_ERR_PARSE,
               "treedraws text does not start with \"ndpost ntree p\"");
    BI_REQUIRE(S64 <= 1 << 24 && m64 <= 1 << 24 && S64 * m64 <= (int64_t)1 << 31, BARTINT_ERR_PARSE,
               "implausible treedraws header (%lld draws x %lld trees)", (long long)S64, (long long)m64);
    BI_REQUIRE(p64 == d, BARTINT_ERR_ARG, "treedraws header says p = %lld but d = %d cut-point lists were given",
               (long long)p64, d);
    f->S = (int32_t)S64; f->m = (int32_t)m64;
    const int64_t n_trees = S64 * m64;
    // pass 1: locate each tree's node lines
    std::vector<const char*> first_line((size_t)n_trees);
    f->tree_off.assign((size_t)n_trees + 1, 0);
    for (int64_t k = 0; k < n_trees; ++k) {
        int64_t nn = 0;
        BI_REQUIRE(parse_i64(p, end, nn) && nn >= 1 && nn < (int64_t)1 << 30, BARTINT_ERR_PARSE,
                   "tree %lld: missing or bad node count", (long long)k);
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        BI_REQUIRE(nl != nullptr, BARTINT_ERR_PARSE, "tree %lld: text ends after the node count", (long long)k);
        p = nl + 1;
        first_line[(size_t)k] = p;
        for (int64_t q = 0; q < nn; ++q) {
            nl = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (nl == nullptr) {
                BI_REQUIRE(q == nn - 1 && p < end, BARTINT_ERR_PARSE, "tree %lld: %lld node lines announced, text ends after %lld",
                           (long long)k, (long long)nn, (long long)q);
                p = end;
            } else {
                p = nl + 1;
            }
        }
        f->tree_off[(size_t)k + 1] = f->tree_off[(size_t)k] + nn;
    }
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\0')) ++p;
    BI_REQUIRE(p == end, BARTINT_ERR_PARSE, "trailing text after the last tree");
    const int64_t nn_total = f->tree_off[(size_t)n_trees];
    f->var.resize((size_t)nn_total); f->cut.resize((size_t)nn_total); f->left.resize((size_t)nn_total);
    f->right.resize((size_t)nn_total); f->mu.resize((size_t)nn_total);
    // pass 2: parse + heap ids -> pre-order
    int64_t bad_tree = -1;
#pragma omp parallel
    {
        struct Node { uint64_t nid; int32_t var, cut; double theta; };
        std::vector<Node> nodes;
        std::vector<std::pair<int64_t, int32_t>> stack;      // (index into nodes, parent slot or -1 - (parent slot) for right)
#pragma omp for schedule(static)
        for (int64_t k = 0; k < n_trees; ++k) {
            const int64_t base = f->tree_off[(size_t)k];
            const int32_t nn = (int32_t)(f->tree_off[(size_t)k + 1] - base);
            const char* q = first_line[(size_t)k];
            nodes.clear();
            bool ok = true;
            for (int32_t a = 0; a < nn && ok; ++a) {
                int64_t nid = 0, v = 0, c = 0;
                ok = parse_i64(q, end, nid) && parse_i64(q, end, v) && parse_i64(q, end, c) && nid >= 1;
                if (!ok) break;
                while (q < end && (*q == ' ' || *q == '\t')) ++q;
                char* after = nullptr;
                const double th = strtod(q, &after);           // text is NUL terminated (ABI contract)
                ok = after != q && after <= end;
                q = after;
                nodes.push_back(Node{(uint64_t)nid, (int32_t)std::min<int64_t>(v, INT32_MAX), (int32_t)std::min<int64_t>(c, INT32_MAX), th});
            }
            if (ok) {
                std::sort(nodes.begin(), nodes.end(), [](const Node& x, const Node& y) { return x.nid < y.nid; });
                for (int32_t a = 1; a < nn; ++a) ok = ok && nodes[(size_t)a].nid != nodes[(size_t)a - 1].nid;
            }
            auto find = [&](uint64_t nid) -> int64_t {
                auto it = std::lower_bound(nodes.begin(), nodes.end(), nid, [](const Node& x, uint64_t id) { return x.nid < id; });
                return (it != nodes.end() && it->nid == nid) ? (int64_t)(it - nodes.begin()) : -1;
            };
            int32_t emitted = 0;
            if (ok) {
                const int64_t root = find(1);
                ok = root >= 0;
                stack.clear();
                if (ok) stack.emplace_back(root, INT32_MIN);
                while (!stack.empty() && ok) {
                    const int64_t a = stack.back().first;
                    const int32_t link = stack.back().second;
                    stack.pop_back();
                    const int32_t me = emitted++;
                    if (link != INT32_MIN && link < 0) f->right[(size_t)(base + (-1 - link))] = me;   // right child of that slot
                    const Node& nd = nodes[(size_t)a];
                    const int64_t l = nd.nid <= (UINT64_MAX >> 1) ? find(2 * nd.nid) : -1;
                    if (l >= 0) {
                        const int64_t r = find(2 * nd.nid + 1);
                        ok = r >= 0 && nd.var >= 0 && nd.var < d &&
                             nd.cut < f->cut_off[(size_t)nd.var + 1] - f->cut_off[(size_t)nd.var];
                        if (!ok) break;
                        f->var[(size_t)(base + me)] = nd.var; f->cut[(size_t)(base + me)] = nd.cut;
                        f->left[(size_t)(base + me)] = me + 1; f->right[(size_t)(base + me)] = -1;
                        f->mu[(size_t)(base + me)] = 0.0;
                        stack.emplace_back(r, -1 - me);
                        stack.emplace_back(l, me);
                    } else {
                        f->var[(size_t)(base + me)] = -1; f->cut[(size_t)(base + me)] = 0;
                        f->left[(size_t)(base + me)] = -1; f->right[(size_t)(base + me)] = -1;
                        f->mu[(size_t)(base + me)] = nd.theta;
                    }
                }
                ok = ok && emitted == nn;                     // every listed node reachable from the root
            }
            if (!ok) {
#pragma omp critical
                if (bad_tree < 0 || k < bad_tree) bad_tree = k;
            }
        }
    }
    BI_REQUIRE(bad_tree < 0, BARTINT_ERR_PARSE,
               "tree %lld (draw %lld, tree %lld): malformed node lines, missing root/child, unreachable node, or a split outside the cut-point lists",
               (long long)bad_tree, (long long)(bad_tree / std::max<int64_t>(m64, 1)), (long long)(bad_tree % std::max<int64_t>(m64, 1)));
    return BARTINT_OK;
}

int canonicalise_soa(int32_t S, int32_t m, int32_t d, const int64_t* off, const int32_t* var, const int32_t* cut,
                     const int32_t* left, const int32_t* right, const double* mu, bartint_forest* f) {
    const int64_t n_trees = (int64_t)S * m;
    BI_REQUIRE(off[0] == 0, BARTINT_ERR_ARG, "tree_node_offset[0] must be 0");
    for (int64_t k = 0; k < n_trees; ++k)
        BI_REQUIRE(off[k + 1] > off[k], BARTINT_ERR_ARG, "tree %lld has no nodes", (long long)k);
    const int64_t nn = off[n_trees];
    f->tree_off.assign(off, off + n_trees + 1);
    f->var.resize((size_t)nn); f->cut.resize((size_t)nn); f->left.resize((size_t)nn); f->right.resize((size_t)nn);
    f->mu.resize((size_t)nn);
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_trees; ++k) {
        const int64_t base = off[k];
        const int32_t cnt = (int32_t)(off[k + 1] - base);
        // pre-order renumbering from the root (node 0 of the tree)
        std::vector<int32_t> order; order.reserve(cnt);
        std::vector<int32_t> newid((size_t)cnt, -1), stack;
        stack.push_back(0);
        bool ok = true;
        while (!stack.empty() && ok) {
            int32_t a = stack.back(); stack.pop_back();
            if (a < 0 || a >= cnt || newid[a] != -1) { ok = false; break; }
            newid[a] = (int32_t)order.size(); order.push_back(a);
            if (var[base + a] >= 0) {
                if (var[base + a] >= d) { ok = false; break; }
                stack.push_back(right[base + a]);
                stack.push_back(left[base + a]);
            }
        }
        if (!ok || (int32_t)order.size() != cnt) {
#pragma omp atomic write
            bad = 1;
            continue;
        }
        for (int32_t b = 0; b < cnt; ++b) {
            int32_t a = order[b];
            bool leaf = var[base + a] < 0;
            f->var[(size_t)(base + b)] = leaf ? -1 : var[base + a];
            f->cut[(size_t)(base + b)] = leaf ? 0 : cut[base + a];
            f->left[(size_t)(base + b)] = leaf ? -1 : newid[left[base + a]];
            f->right[(size_t)(base + b)] = leaf ? -1 : newid[right[base + a]];
            f->mu[(size_t)(base + b)] = leaf ? mu[base + a] : 0.0;
        }
    }
    BI_REQUIRE(!bad, BARTINT_ERR_ARG, "SoA forest is not a set of binary trees rooted at node 0 (bad child index, cycle, unreachable node or variable >= d)");
    return BARTINT_OK;
}

// Forests whose splits are given as VALUES (dbarts >= 0.9-9 `fit$getTrees()` / `extract(fit, "trees")` data frames,
// bartMachine's raw node data): the cut-point lists are built from the split values themselves -- per variable the
// sorted distinct values used anywhere in the posterior -- so routing on the resulting codes is exact for the
// package's own comparison (x <= c or x < c) whatever grid the package drew its cuts from.  left / right may be NULL:
// the nodes of a tree are then in pre-order (a node's left child is the next node).  Leaf values are multiplied by
// leaf_scale.  Fills f's canonical SoA and cut lists.
int value_splits_to_soa(int32_t S, int32_t m, int32_t d, const int64_t* off, const int32_t* var, const double* value,
                        const int32_t* left, const int32_t* right, double leaf_scale, bartint_forest* f) {
    const int64_t n_trees = (int64_t)S * m;
    BI_REQUIRE(off[0] == 0, BARTINT_ERR_ARG, "tree_node_offset[0] must be 0");
    for (int64_t k = 0; k < n_trees; ++k)
        BI_REQUIRE(off[k + 1] > off[k], BARTINT_ERR_ARG, "tree %lld has no nodes", (long long)k);
    const int64_t nn = off[n_trees];
    // 1. per variable: sorted distinct split values
    std::vector<std::vector<double>> cuts((size_t)d);
    for (int64_t g = 0; g < nn; ++g) {
        if (var[g] < 0) continue;
        BI_REQUIRE(var[g] < d, BARTINT_ERR_ARG, "node %lld splits on variable %d but the model has %d", (long long)g, var[g], d);
        BI_REQUIRE(value[g] == value[g], BARTINT_ERR_ARG, "node %lld has a NaN split value", (long long)g);
        cuts[(size_t)var[g]].push_back(value[g]);
    }
    f->cut_off.assign((size_t)d + 1, 0);
    f->cut_val.clear();
    for (int32_t j = 0; j < d; ++j) {
        std::vector<double>& c = cuts[(size_t)j];
        std::sort(c.begin(), c.end());
        c.erase(std::unique(c.begin(), c.end()), c.end());
        BI_REQUIRE(c.size() <= 255, BARTINT_ERR_UNSUPPORTED,
                   "variable %d is split at %zu distinct values; codes are uint8 (at most 255 cut points per variable)", j, c.size());
        f->cut_val.insert(f->cut_val.end(), c.begin(), c.end());
        f->cut_off[(size_t)j + 1] = (int32_t)f->cut_val.size();
    }
    // 2. children (explicit, or implied by pre-order) and cut indices
    std::vector<int32_t> cut((size_t)nn, 0), lft, rgt;
    std::vector<double> mu((size_t)nn, 0.0);
    const bool implied = left == nullptr || right == nullptr;
    if (implied) { lft.assign((size_t)nn, -1); rgt.assign((size_t)nn, -1); }
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_trees; ++k) {
        const int64_t base = off[k];
        const int32_t cnt = (int32_t)(off[k + 1] - base);
        if (implied) {                      // pre-order: the right child of a node follows its whole left subtree
            std::vector<int32_t> pend;      // internal nodes waiting for their right child
            for (int32_t a = 0; a < cnt; ++a) {
                if (var[base + a] >= 0) {
                    if (a + 1 >= cnt) { bad = 1; break; }
                    lft[(size_t)(base + a)] = a + 1;
                    pend.push_back(a);
                } else if (!pend.empty()) {
                    if (a + 1 >= cnt) { bad = 1; break; }
                    rgt[(size_t)(base + pend.back())] = a + 1;
                    pend.pop_back();
                } else if (a + 1 != cnt) {
                    bad = 1;                // nodes after the tree is complete
                    break;
                }
            }
            if (!pend.empty()) bad = 1;
        }
        for (int32_t a = 0; a < cnt; ++a) {
            const int64_t g = base + a;
            if (var[g] >= 0) {
                const std::vector<double>& c = cuts[(size_t)var[g]];
                cut[(size_t)g] = (int32_t)(std::lower_bound(c.begin(), c.end(), value[g]) - c.begin());
            } else {
                mu[(size_t)g] = leaf_scale * value[g];
            }
        }
    }
    BI_REQUIRE(!bad, BARTINT_ERR_ARG, "node arrays are not complete binary trees in pre-order");
    return canonicalise_soa(S, m, d, off, var, cut.data(), implied ? lft.data() : left, implied ? rgt.data() : right, mu.data(), f);
}

static int forest_check_cuts(bartint_forest* f);
static int forest_finish_host_trees(bartint_forest* f, PhaseTimer& pt);

int forest_finish_host(bartint_forest* f) {
    PhaseTimer pt;
    const int64_t n_trees = (int64_t)f->S * f->m;
    f->n_nodes = f->tree_off[(size_t)n_trees];
    BI_REQUIRE(f->n_nodes < (int64_t)1 << 31, BARTINT_ERR_UNSUPPORTED, "forest has more than 2^31 nodes");
    int rc_cuts = forest_check_cuts(f);
    if (rc_cuts) return rc_cuts;
    return forest_finish_host_trees(f, pt);
}

// cut-point lists: ascending, at most 255 per variable (codes are uint8); fills f->max_cuts
static int forest_check_cuts(bartint_forest* f) {
    f->max_cuts = 0;
    for (int32_t j = 0; j < f->d; ++j) {
        int nc = f->cut_off[(size_t)j + 1] - f->cut_off[(size_t)j];
        BI_REQUIRE(nc >= 0, BARTINT_ERR_ARG, "cut_offsets must be non-decreasing");
        f->max_cuts = std::max(f->max_cuts, nc);
        for (int k = 1; k < nc; ++k)
            BI_REQUIRE(f->cut_val[(size_t)f->cut_off[(size_t)j] + k] >= f->cut_val[(size_t)f->cut_off[(size_t)j] + k - 1],
                       BARTINT_ERR_ARG, "cut points of variable %d are not ascending", j);
    }
    BI_REQUIRE(f->max_cuts <= 255, BARTINT_ERR_UNSUPPORTED, "more than 255 cut points for one variable (codes are uint8)");
    BI_REQUIRE(f->d < 65535, BARTINT_ERR_UNSUPPORTED, "more than 65534 variables");
    return BARTINT_OK;
}

static int forest_finish_host_trees(bartint_forest* f, PhaseTimer& pt) {
    const int64_t n_trees = (int64_t)f->S * f->m;
    int64_t n_leaves = 0;
    int32_t max_depth = 0, max_internal = 0;
    int bad = 0;
    // internal nodes by depth class (0, 1, 2, 3, >= 4) and the mask rows they reference in the bit-sliced kernel
    // (depth + 1 for depth >= 1): the kernel's shared-memory floor, reported by bartint_forest_bitslice_info
    int64_t cls0 = 0, cls1 = 0, cls2 = 0, cls3 = 0, cls4 = 0, refs = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_leaves, cls0, cls1, cls2, cls3, cls4, refs) reduction(max : max_depth, max_internal)
    for (int64_t k = 0; k < n_trees; ++k) {
        const int64_t base = f->tree_off[(size_t)k];
        const int32_t cnt = (int32_t)(f->tree_off[(size_t)k + 1] - base);
        int32_t internal = 0;
        // depth via pre-order scan with a stack of "subtree ends"
        std::vector<int32_t> depth((size_t)cnt, 0);
        for (int32_t a = 0; a < cnt; ++a) {
            if (f->var[(size_t)(base + a)] >= 0) {
                ++internal;
                const int32_t dp = depth[(size_t)a];
                if (dp == 0) ++cls0; else if (dp == 1) ++cls1; else if (dp == 2) ++cls2; else if (dp == 3) ++cls3; else ++cls4;
                if (dp >= 1) refs += dp + 1;
                int32_t v = f->var[(size_t)(base + a)], c = f->cut[(size_t)(base + a)];
                if (c < 0 || c >= f->cut_off[(size_t)v + 1] - f->cut_off[(size_t)v]) bad = 1;
                depth[(size_t)f->left[(size_t)(base + a)]] = depth[(size_t)a] + 1;
                depth[(size_t)f->right[(size_t)(base + a)]] = depth[(size_t)a] + 1;
            } else {
                ++n_leaves;
            }
            max_depth = std::max(max_depth, depth[(size_t)a]);
        }
        if (cnt != 2 * internal + 1) bad = 1;
        max_internal = std::max(max_internal, internal);
    }
    BI_REQUIRE(!bad, BARTINT_ERR_ARG, "forest has a cut index outside its variable's cut-point list or a malformed tree");
    f->n_leaves = n_leaves;
    f->max_depth = max_depth;
    f->max_internal = max_internal;
    f->bs_class_nodes[0] = cls0; f->bs_class_nodes[1] = cls1; f->bs_class_nodes[2] = cls2; f->bs_class_nodes[3] = cls3;
    f->bs_class_nodes[4] = cls4; f->bs_refs_total = refs;
    // for the bit-sliced kernel: the longest draw and the largest |leaf value| (fixed-point scale)
    int32_t draw_internal_max = 0;
    double max_abs_mu = 0.0;
    bool mu_finite = true;
#pragma omp parallel for schedule(static) reduction(max : draw_internal_max, max_abs_mu) reduction(&& : mu_finite)
    for (int32_t s = 0; s < f->S; ++s) {
        const int64_t g0 = f->tree_off[(size_t)s * (size_t)f->m], g1 = f->tree_off[(size_t)(s + 1) * (size_t)f->m];
        draw_internal_max = std::max(draw_internal_max, (int32_t)((g1 - g0 - f->m) / 2));
        for (int64_t g = g0; g < g1; ++g)
            if (f->var[(size_t)g] < 0) {
                const double a = std::fabs(f->mu[(size_t)g]);
                if (!(a <= 1.7e308)) mu_finite = false;
                else max_abs_mu = std::max(max_abs_mu, a);
            }
    }
    f->draw_internal_max = draw_internal_max;
    f->max_abs_mu = mu_finite ? max_abs_mu : INFINITY;
    bs_configure(f);
    pt.lap("validate + stats");
    return BARTINT_OK;
}

// ---- page-locked staging blocks (see StageAlloc in common.cuh) ----
namespace {
struct StageBlock { void* p; size_t bytes; bool pinned; cudaEvent_t busy; };
struct StageCache {
    std::vector<StageBlock> free_blocks;
    ~StageCache() {
        for (StageBlock& b : free_blocks) {
            if (b.busy) { cudaEventSynchronize(b.busy); cudaEventDestroy(b.busy); }
            if (b.pinned) cudaFreeHost(b.p); else free(b.p);
        }
        (void)cudaGetLastError();
    }
};
thread_local StageCache g_stage;
std::atomic<int> g_stage_pinned{-1};      // -1 unknown, 0 no device (heap), 1 page-locked
constexpr size_t kStageHeader = 64;       // room in front of the user pointer for {bytes, pinned}
}  // namespace

void* pinned_stage_alloc(size_t bytes) {
    const size_t need = ((std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095) + kStageHeader;
    // the SMALLEST cached block that is large enough and whose last copy has completed (first fit let a 100 KB request
    // take the 700 KB block the next request needed, which then cost a cudaHostAlloc in every step)
    int best = -1;
    for (size_t k = 0; k < g_stage.free_blocks.size(); ++k) {
        StageBlock& b = g_stage.free_blocks[k];
        if (b.bytes < need || b.bytes > 4 * need + (1u << 20)) continue;
        if (best >= 0 && g_stage.free_blocks[(size_t)best].bytes <= b.bytes) continue;
        if (b.busy) {
            if (cudaEventQuery(b.busy) != cudaSuccess) { (void)cudaGetLastError(); continue; }
            cudaEventDestroy(b.busy);
            b.busy = nullptr;
        }
        best = (int)k;
    }
    if (best >= 0) {
        StageBlock got = g_stage.free_blocks[(size_t)best];
        g_stage.free_blocks.erase(g_stage.free_blocks.begin() + best);
        size_t* hdr = static_cast<size_t*>(got.p);
        hdr[0] = got.bytes; hdr[1] = got.pinned ? 1 : 0;
        return static_cast<char*>(got.p) + kStageHeader;
    }
    void* p = nullptr;
    bool pinned = false;
    if (g_stage_pinned.load() != 0) {
        int n_dev = 0;
        if (g_stage_pinned.load() < 0) g_stage_pinned.store(cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0 ? 1 : 0);
        (void)cudaGetLastError();
        if (g_stage_pinned.load() == 1 && cudaHostAlloc(&p, need, cudaHostAllocPortable) == cudaSuccess) pinned = true;
        else { (void)cudaGetLastError(); p = nullptr; }
    }
    if (!p) p = malloc(need);
    if (!p) throw std::bad_alloc();
    size_t* hdr = static_cast<size_t*>(p);
    hdr[0] = need; hdr[1] = pinned ? 1 : 0;
    return static_cast<char*>(p) + kStageHeader;
}

void pinned_stage_free(void* user, size_t) {
    if (!user) return;
    void* p = static_cast<char*>(user) - kStageHeader;
    const size_t* hdr = static_cast<const size_t*>(p);
    StageBlock b{p, hdr[0], hdr[1] != 0, nullptr};
    if (b.pinned) {      // a copy out of this block may still be queued on the upload (default) stream
        if (cudaEventCreateWithFlags(&b.busy, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(b.busy, nullptr) != cudaSuccess) {
            (void)cudaGetLastError();
            cudaStreamSynchronize(nullptr);
            if (b.busy) { cudaEventDestroy(b.busy); b.busy = nullptr; }
        }
    }
    if (g_stage.free_blocks.size() >= 32) {          // bound the cache: drop the oldest block
        StageBlock old = g_stage.free_blocks.front();
        g_stage.free_blocks.erase(g_stage.free_blocks.begin());
        if (old.busy) { cudaEventSynchronize(old.busy); cudaEventDestroy(old.busy); }
        if (old.pinned) cudaFreeHost(old.p); else free(old.p);
    }
    g_stage.free_blocks.push_back(b);
}

template <typename T, typename A>
static int upload(T** dptr, const std::vector<T, A>& h) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    // stream-ordered allocator on the default stream: cached blocks make a per-step forest rebuild cheap
    BI_CUDA(cudaMallocAsync((void**)dptr, bytes, nullptr));
    if (!h.empty()) BI_CUDA(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, nullptr));
    return BARTINT_OK;
}

static int table_bits_from_env() {
    const char* e = getenv("BARTINT_TABLE_BITS");
    int nb = e ? atoi(e) : 5;
    if (nb < 4) nb = 4;
    if (nb > 8) nb = 8;
    return nb;
}

// Chooses the fast kernel (f->table_ok / gather / nb / g_regs / count_mode), packs the trees of every draw into
// group records and fills f->n_groups / f->draw_group_off.  tree_off: first node of each tree (int32 copy).
// The trees of one table: a table indexes <= 8 predicate bits and every tree in it has at least one split, so eight
// slots always suffice (no heap traffic in the packing loops, which run once per draw and candidate K).
struct TreeList {
    int32_t v[8];
    int n = 0;
    void push_back(int32_t t) { v[n++] = t; }
    bool empty() const { return n == 0; }
    const int32_t* begin() const { return v; }
    const int32_t* end() const { return v + n; }
};

void plan_fast_path(bartint_forest* f, const int32_t* tree_off, FastPlan& P) {
    PhaseTimer pt;
    const int64_t n_trees = (int64_t)f->S * f->m;
    const int64_t nn = f->n_nodes;
    // ---- fast-path planning: greedy packing of consecutive trees of a draw into groups of <= NB predicate bits ----
    f->nb = table_bits_from_env();
    // both fast kernels need 7-bit codes (byte-wise compare); the table kernel keeps the code rows of one CTA tile in
    // shared memory, the gather kernel keeps a point's codes in <= 4 registers
    f->table_ok = n_trees > 0 && f->max_cuts <= 127 && (int64_t)f->d * TK_TILE_MIN <= 160 * 1024;
    const char* kenv = getenv("BARTINT_KERNEL");
    f->gather = f->table_ok && f->d <= 16 && f->d >= 1 && !(kenv && strcmp(kenv, "table") == 0);
    if (f->gather && f->nb > 5) f->nb = 5;                 // slot weights -(8 << bit) must fit a signed byte
    const char* cenv = getenv("BARTINT_COUNT_SPLITS");
    // opt-in: trees of <= 1 (or 2) splits are counted, not evaluated, in sums-only evaluations
    f->count_level = (f->gather && cenv && (cenv[0] == '1' || cenv[0] == '2')) ? cenv[0] - '0' : 0;
    f->count_mode = f->count_level > 0;
    f->g_regs = tg_regs_for(f->d);
    f->n_groups = 0;
    f->draw_group_off.clear();
    if (!f->table_ok) return;
    const int NB = f->nb;
    const int DW = P.DW = f->gather ? TG_DESC / 4 : 2 * NB;       // descriptor words per group
    std::vector<uint8_t>& node_bit = P.node_bit;
    node_bit.assign((size_t)nn, 0);
    // Bit order inside a group: table lookups of lanes whose indices differ only in a HIGH bit hit the same bank pair
    // (a 2^NB x 8 B table wraps the 32 banks 2^(NB-4) times), lookups with equal indices are a broadcast.  The
    // predicates least likely to differ between points -- cuts nearest the edge of their variable's cut grid -- are
    // therefore given the highest bits.  Pure layout choice: the builder and the kernels read node_bit / desc.
    auto extremeness = [&](int64_t g) {
        const int v = f->var[(size_t)g];
        const double nc = (double)(f->cut_off[(size_t)v + 1] - f->cut_off[(size_t)v]);
        return std::fabs(2.0 * (f->cut[(size_t)g] + 1.0) / (nc + 1.0) - 1.0);
    };
    // gather format: round 0 can address variables [0, 8), round 1 variables [base1, base1 + 8); 4 slots per round
    const int base1 = 4 * (f->g_regs - 2);
    struct Slot { int64_t node; int round, pos; };
    // draws are independent: each is planned on its own (OpenMP) into a DrawPlan, then the plans are concatenated
    struct DrawPlan {
        std::vector<int32_t> group_tree_off;   // first tree of each group of the draw
        std::vector<uint32_t> hdr;             // flags per group
        std::vector<uint32_t> split;           // gather DUAL records: number of trees in table 1
        std::vector<uint32_t> desc;            // DW words per group
    };
    std::vector<int32_t>& tree_perm = P.tree_perm;             // permuted position -> tree id (identity: table format)
    tree_perm.resize((size_t)n_trees);
    for (int64_t k = 0; k < n_trees; ++k) tree_perm[(size_t)k] = (int32_t)k;
    std::vector<DrawPlan> plans((size_t)f->S);
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < f->S; ++s) {
        std::vector<int32_t>& group_tree_off = plans[(size_t)s].group_tree_off;
        std::vector<uint32_t>& hdr = plans[(size_t)s].hdr;
        std::vector<uint32_t>& split = plans[(size_t)s].split;
        std::vector<uint32_t>& desc = plans[(size_t)s].desc;
        {
            const size_t guess = (size_t)f->m / 4 + 4;               // a draw's records: one allocation per array
            group_tree_off.reserve(guess); hdr.reserve(guess); split.reserve(guess); desc.reserve(guess * (size_t)DW);
        }
        auto open_group = [&](int64_t k, uint32_t flags) {
            group_tree_off.push_back((int32_t)k);
            hdr.push_back(flags);
            split.push_back(0u);
            desc.resize(desc.size() + (size_t)DW, 0u);           // unused slots: bias 0 -> predicate 0, weight 0
        };
        auto internal_nodes = [&](int64_t k, std::vector<int64_t>& out) {
            out.clear();
            const int32_t base = tree_off[(size_t)k], cnt = tree_off[(size_t)k + 1] - base;
            for (int32_t a = 0; a < cnt; ++a)
                if (f->var[(size_t)(base + a)] >= 0) out.push_back((int64_t)base + a);
        };

        if (!f->gather) {
            // ---------------- table format: up to NB nodes per group, any variables ----------------
            static thread_local std::vector<int64_t> gnodes, tn;
            gnodes.clear(); tn.clear();
            auto close_group = [&]() {
                if (gnodes.empty()) return;
                std::stable_sort(gnodes.begin(), gnodes.end(),
                                 [&](int64_t a, int64_t b2) { return extremeness(a) < extremeness(b2); });
                uint32_t* gd = &desc[desc.size() - (size_t)DW];
                for (size_t b2 = 0; b2 < gnodes.size(); ++b2) {
                    const size_t g = (size_t)gnodes[b2];
                    node_bit[g] = (uint8_t)b2;
                    gd[2 * b2] = (uint32_t)f->var[g];
                    gd[2 * b2 + 1] = (uint32_t)(127 - f->cut[g]) * 0x01010101u;
                }
                gnodes.clear();
            };
            // The trees of a draw may be summed in any order, so they are bin-packed into the groups (largest trees
            // first, fullest group that still takes the tree) and the draw's trees are PERMUTED: group g covers
            // permuted positions [group_tree_off[g], group_tree_off[g+1]), tree_perm maps to tree ids.  (Consecutive
            // greedy packing left ~10 % of the predicate bits unused at cfg 3.)
            struct TBin { int nodes = 0; TreeList trees; };
            // scratch that lives as long as the (pooled) host thread: no allocation per draw
            static thread_local std::vector<TBin> bins;
            static thread_local std::vector<std::pair<int, int32_t>> order;          // (internal nodes, tree)
            static thread_local std::vector<int32_t> stumps, walk_trees;
            bins.clear(); order.clear(); stumps.clear(); walk_trees.clear();
            for (int32_t t = 0; t < f->m; ++t) {
                internal_nodes((int64_t)s * f->m + t, tn);
                if (tn.empty()) stumps.push_back(t);
                else if ((int)tn.size() > NB) walk_trees.push_back(t);
                else order.emplace_back((int)tn.size(), t);
            }
            std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b2) { return a.first > b2.first; });
            for (const auto& kt : order) {
                TBin* best = nullptr;
                for (TBin& b : bins)
                    if (b.nodes + kt.first <= NB && (best == nullptr || b.nodes > best->nodes)) best = &b;
                if (best == nullptr) { bins.emplace_back(); best = &bins.back(); }
                best->nodes += kt.first;
                best->trees.push_back(kt.second);
            }
            int32_t* perm = tree_perm.data() + (size_t)s * f->m;
            int32_t pos = 0;
            if (bins.empty() && !stumps.empty()) bins.emplace_back();      // a draw of stumps (and WALK trees) only
            for (size_t bi = 0; bi < bins.size(); ++bi) {
                open_group((int64_t)s * f->m + pos, 0u);
                if (bi == 0)                                      // stumps only add a constant to the table
                    for (int32_t t : stumps) perm[pos++] = (int32_t)((int64_t)s * f->m + t);
                for (int32_t t : bins[bi].trees) {
                    perm[pos++] = (int32_t)((int64_t)s * f->m + t);
                    internal_nodes((int64_t)s * f->m + t, tn);
                    gnodes.insert(gnodes.end(), tn.begin(), tn.end());
                }
                close_group();
            }
            for (int32_t t : walk_trees) {                        // WALK record holding just this tree
                open_group((int64_t)s * f->m + pos, TK_FLAG_WALK);
                perm[pos++] = (int32_t)((int64_t)s * f->m + t);
            }
            hdr.back() |= TK_FLAG_END_DRAW;
            continue;
        }

        // ---------------- gather format ----------------
        // DUAL record (the common kind): two independent 16-entry tables.  Table 1 holds trees whose <= 4 nodes use
        // variables of code words (0,1); table 2 gathers from words (R-2,R-1) [pairsel 0], (0,R-1) [pairsel 1] or,
        // with one more PRMT per point, (0,1,R-1) [pairsel 2].  One PRMT gather + one IDP.4A per table and point.
        // The trees of a draw may be summed in any order, so they are bin-packed into the tables and the draw's
        // trees are PERMUTED: record g covers permuted positions [group_tree_off[g], group_tree_off[g+1]), tree_perm
        // maps to tree ids.  Packing: for K = ceil(nodes / 8), K + 1, ... try to fit every tree into K table-1 bins
        // and K table-2 bins (largest trees first; fullest bin that takes the tree, never widening a table-2 bin to
        // three words when a narrower home exists); the first K that fits is almost always the lower bound.
        // JOINT record: one 2^NB table over two gather rounds, for 5-node trees.  WALK: no table.
        const int R = f->g_regs;
        const unsigned last = 1u << (R - 1);
        const unsigned m_side1 = 3u, m_pair0 = (1u << (R - 2)) | last, m_pair1 = 1u | last, m_triple = 3u | last;
        auto within = [](unsigned regs, unsigned mask) { return (regs & ~mask) == 0u; };
        auto pair_ok = [&](unsigned regs) { return within(regs, m_pair0) || within(regs, m_pair1); };
        auto side2_ok = [&](unsigned regs) { return pair_ok(regs) || within(regs, m_triple); };
        struct Item { int k; unsigned regs; int flex; int32_t tree; };
        struct Bin { int nodes = 0; unsigned regs = 0; TreeList trees; };
        // scratch that lives as long as the (pooled) host thread: no allocation per draw
        static thread_local std::vector<Item> items;
        static thread_local std::vector<Bin> bins1, bins2;
        static thread_local std::vector<int32_t> stumps, joint_trees, walk_trees;
        static thread_local std::vector<int64_t> tn;
        static thread_local std::vector<Item> countable;         // counting mode: trees of <= count_level splits
        items.clear(); stumps.clear(); joint_trees.clear(); walk_trees.clear(); tn.clear(); countable.clear();
        for (int32_t t = 0; t < f->m; ++t) {
            const int64_t kt = (int64_t)s * f->m + t;
            const int32_t nb0 = tree_off[(size_t)kt], ne0 = tree_off[(size_t)kt + 1];
            if (ne0 - nb0 == 1) { stumps.push_back(t); continue; }
            unsigned regs = 0;
            for (int32_t g = nb0; g < ne0; ++g)
                if (f->var[(size_t)g] >= 0) regs |= 1u << (f->var[(size_t)g] >> 2);
            const int k = (ne0 - nb0 - 1) / 2;            // internal nodes of a binary tree with ne0 - nb0 nodes
            const bool dual_able = within(regs, m_side1) || side2_ok(regs);      // == tg_dual_able(regs, R)
            // counted, not evaluated, when only sums are wanted.  Only trees that fit a DUAL table: an evaluation that
            // wants per-point results evaluates the COUNTABLE records like any other (a two-split tree on words (0,2)
            // or (1,2) of a four-word point stays in the evaluated part; count_draw_sums_kernel applies the same test)
            if (f->count_mode && k <= f->count_level && dual_able) {
                countable.push_back(Item{k, regs, (int)within(regs, m_side1) + (int)pair_ok(regs), t});
                continue;
            }
            if (k <= 4 && dual_able) {
                items.push_back(Item{k, regs, (int)within(regs, m_side1) + (int)pair_ok(regs), t});
            } else {
                (k <= NB ? joint_trees : walk_trees).push_back(t);
            }
        }
        const std::vector<Item>* cur_items = &items;             // the set try_pack / pack_and_emit work on
        auto try_pack = [&](int K) -> bool {
            bins1.assign((size_t)K, Bin{});
            bins2.assign((size_t)K, Bin{});
            size_t lo1 = 0, lo2 = 0;                      // bins before these are full: never looked at again
            for (const Item& it : *cur_items) {
                Bin* best = nullptr;
                int best_key = 0;                         // lower is better: widening << 8 | (4 - fill) << 1 | side 2
                const int exact1 = it.k << 1, exact2 = (it.k << 1) | 1;      // keys of bins this tree would fill exactly
                if (within(it.regs, m_side1))
                    for (size_t q = lo1; q < (size_t)K; ++q) {
                        Bin& b = bins1[q];
                        if (b.nodes + it.k > 4) continue;
                        const int key = (4 - b.nodes) << 1;
                        if (best == nullptr || key < best_key) { best = &b; best_key = key; }
                        if (key == exact1) break;                           // nothing beats an exact fit on side 1
                    }
                if (best_key != exact1 || best == nullptr)
                    for (size_t q = lo2; q < (size_t)K; ++q) {
                        Bin& b = bins2[q];
                        const unsigned u = b.regs | it.regs;
                        if (b.nodes + it.k > 4 || !side2_ok(u)) continue;
                        const int widen = (!pair_ok(u) && pair_ok(b.regs)) ? 1 : 0;
                        const int key = (widen << 8) | ((4 - b.nodes) << 1) | 1;
                        if (best == nullptr || key < best_key) { best = &b; best_key = key; }
                        if (key == exact2) break;                           // best possible key on side 2
                    }
                if (best == nullptr) return false;
                best->nodes += it.k;
                best->regs |= it.regs;
                best->trees.push_back(it.tree);
                while (lo1 < (size_t)K && bins1[lo1].nodes >= 4) ++lo1;
                while (lo2 < (size_t)K && bins2[lo2].nodes >= 4) ++lo2;
            }
            return true;
        };
        int32_t* perm = tree_perm.data() + (size_t)s * f->m;
        int32_t pos = 0;                                  // next permuted position inside the draw
        auto emit_table = [&](const Bin& bin, int table, uint32_t* gd) {
            if (bin.trees.empty()) return;
            const int ps = table == 0 ? -1 : within(bin.regs, m_pair0) ? 0 : within(bin.regs, m_pair1) ? 1 : 2;
            uint32_t sel = 0, sel_t = 0, bias = 0, wgt = 0;
            int n = 0;
            for (int32_t t : bin.trees) {
                const int64_t kt = (int64_t)s * f->m + t;
                perm[pos++] = (int32_t)kt;
                for (int32_t gi = tree_off[(size_t)kt], ge = tree_off[(size_t)kt + 1]; gi < ge; ++gi) {
                    const size_t g = (size_t)gi;
                    const int v = f->var[g];
                    if (v < 0) continue;                  // leaf
                    // byte of variable v inside the table's gather operands (see dual_offsets in kernels_gather.cu)
                    int byte = v;                                                   // table 1: words (0,1)
                    if (ps == 0) byte = v - base1;                                  // words (R-2,R-1)
                    else if (ps == 1) byte = v < 4 ? v : v - base1;                 // words (0,R-1)
                    else if (ps == 2) {                                             // (PRMT(w0,w1,sel_t), R-1)
                        if (v < 8) { sel_t |= (uint32_t)v << (4 * n); byte = n; }
                        else byte = v - base1;
                    }
                    node_bit[g] = (uint8_t)n;
                    sel |= (uint32_t)byte << (4 * n);
                    bias |= (uint32_t)(127 - f->cut[g]) << (8 * n);
                    wgt |= ((uint32_t)(-(8 << n)) & 0xFFu) << (8 * n);
                    ++n;
                }
            }
            if (table == 0) { gd[0] |= sel; gd[1] = bias; gd[3] = wgt; }
            else { gd[0] |= sel << 16; gd[2] = bias; gd[4] = wgt; gd[5] = (uint32_t)ps; gd[6] = sel_t; }
        };
        // pack a set of DUAL-able trees into the fewest records and emit them with the given record flags
        auto pack_and_emit = [&](std::vector<Item>& set, uint32_t rec_flags, bool evaluated_part) {
            // stable order: larger trees first, less flexible ones first among equals (insertion sort: <= m items,
            // no temporary buffer)
            for (size_t q = 1; q < set.size(); ++q) {
                const Item it = set[q];
                size_t r = q;
                while (r > 0 && (set[r - 1].k != it.k ? set[r - 1].k < it.k : set[r - 1].flex > it.flex)) {
                    set[r] = set[r - 1];
                    --r;
                }
                set[r] = it;
            }
            cur_items = &set;
            int set_nodes = 0;
            for (const Item& it : set) set_nodes += it.k;
            // Every tree of the set fits an empty table on its own (that is what put it into the set), so
            // K = number of trees always succeeds and the search ends there at the latest.
            const int K_max = std::max<int>(1, (int)set.size());
            int K = std::min(K_max, std::max(1, (set_nodes + 7) / 8));
            bool packed = try_pack(K);
            while (!packed && K < K_max) packed = try_pack(++K);
            if (!packed) {                                // unreachable by the argument above; never lose a tree
                K = K_max;
                bins1.assign((size_t)K, Bin{});
                bins2.assign((size_t)K, Bin{});
                for (size_t q = 0; q < set.size(); ++q) {
                    Bin& b = within(set[q].regs, m_side1) ? bins1[q] : bins2[q];
                    b.nodes = set[q].k;
                    b.regs = set[q].regs;
                    b.trees.push_back(set[q].tree);
                }
            }
            // emit, sorted by table-2 kind so the kernel's per-record variant branch changes rarely
            static thread_local std::vector<int> rec_order;
            rec_order.clear();
            for (int r = 0; r < K; ++r)
                if (!bins1[(size_t)r].trees.empty() || !bins2[(size_t)r].trees.empty()) rec_order.push_back(r);
            auto kind_of = [&](int r) {
                const Bin& b = bins2[(size_t)r];
                return b.trees.empty() || within(b.regs, m_pair0) ? 0 : within(b.regs, m_pair1) ? 1 : 2;
            };
            std::stable_sort(rec_order.begin(), rec_order.end(), [&](int a, int b2) { return kind_of(a) < kind_of(b2); });
            if (evaluated_part && rec_order.empty() && !stumps.empty() && joint_trees.empty() && walk_trees.empty()) {
                bins1.assign(1, Bin{});
                bins2.assign(1, Bin{});
                rec_order.push_back(0);                   // a draw of stumps only: one record holding the constant
            }
            for (size_t q = 0; q < rec_order.size(); ++q) {
                const int r = rec_order[q];
                open_group((int64_t)s * f->m + pos, rec_flags);
                uint32_t* gd = &desc[desc.size() - (size_t)DW];
                const int32_t start = pos;
                if (evaluated_part && q == 0) {           // stumps only add a constant: all into the first table 1
                    for (int32_t t : stumps) perm[pos++] = (int32_t)((int64_t)s * f->m + t);
                    stumps.clear();
                }
                emit_table(bins1[(size_t)r], 0, gd);
                split.back() = (uint32_t)(pos - start);
                emit_table(bins2[(size_t)r], 1, gd);
            }
        };
        pack_and_emit(items, 0u, true);
        // JOINT records: greedy packing of the remaining small trees, two rounds of <= 4 slots, <= NB bits
        static thread_local std::vector<Slot> gslots;
        gslots.clear();
        int used[2] = {0, 0}, jbits = 0;
        bool jopen = false;
        auto close_joint = [&]() {
            if (!jopen) return;
            std::stable_sort(gslots.begin(), gslots.end(),
                             [&](const Slot& a, const Slot& b2) { return extremeness(a.node) < extremeness(b2.node); });
            uint32_t sel[2] = {0, 0}, bias[2] = {0, 0}, wgt[2] = {0, 0};
            for (size_t b2 = 0; b2 < gslots.size(); ++b2) {
                const Slot& sl = gslots[b2];
                const size_t g = (size_t)sl.node;
                node_bit[g] = (uint8_t)b2;
                sel[sl.round] |= (uint32_t)(f->var[g] - (sl.round ? base1 : 0)) << (4 * sl.pos);
                bias[sl.round] |= (uint32_t)(127 - f->cut[g]) << (8 * sl.pos);
                wgt[sl.round] |= ((uint32_t)(-(8 << b2)) & 0xFFu) << (8 * sl.pos);
            }
            uint32_t* gd = &desc[desc.size() - (size_t)DW];
            gd[0] = sel[0] | (sel[1] << 16); gd[1] = bias[0]; gd[2] = bias[1]; gd[3] = wgt[0]; gd[4] = wgt[1];
            gslots.clear();
            used[0] = used[1] = 0; jbits = 0;
            jopen = false;
        };
        auto place_joint = [&]() -> bool {               // nodes in tn; all or nothing
            const size_t mark = gslots.size();
            int u[2] = {used[0], used[1]};
            int nbits = jbits;
            for (int64_t gi : tn) {
                const int v = f->var[(size_t)gi];
                const bool in0 = v < 8, in1 = v >= base1 && v < base1 + 8;
                int r = -1;
                if (in0 && in1) r = (u[0] <= u[1]) ? 0 : 1;
                else if (in0) r = 0;
                else if (in1) r = 1;
                if (r >= 0 && u[r] >= 4) r = (r == 0 ? (in1 && u[1] < 4 ? 1 : -1) : (in0 && u[0] < 4 ? 0 : -1));
                if (r < 0 || nbits >= NB) { gslots.resize(mark); return false; }
                gslots.push_back(Slot{gi, r, u[r]});
                ++u[r];
                ++nbits;
            }
            used[0] = u[0]; used[1] = u[1]; jbits = nbits;
            return true;
        };
        for (int32_t t : joint_trees) {
            internal_nodes((int64_t)s * f->m + t, tn);
            if (jopen && place_joint()) { perm[pos++] = (int32_t)((int64_t)s * f->m + t); continue; }
            close_joint();
            if (place_joint()) {
                open_group((int64_t)s * f->m + pos, TK_FLAG_JOINT);
                jopen = true;
                if (!stumps.empty()) { for (int32_t st : stumps) perm[pos++] = (int32_t)((int64_t)s * f->m + st); stumps.clear(); }
                perm[pos++] = (int32_t)((int64_t)s * f->m + t);
            } else {
                walk_trees.push_back(t);                  // cannot be slotted (e.g. 5 nodes on variables of one round)
            }
        }
        close_joint();
        for (int32_t t : walk_trees) {                    // one header-only record per tree, walked from global memory
            open_group((int64_t)s * f->m + pos, TK_FLAG_WALK);
            perm[pos++] = (int32_t)((int64_t)s * f->m + t);
        }
        if (!stumps.empty()) {                            // a draw made only of stumps and WALK trees
            open_group((int64_t)s * f->m + pos, 0u);
            for (int32_t t : stumps) perm[pos++] = (int32_t)((int64_t)s * f->m + t);
            split.back() = (uint32_t)stumps.size();
        }
        if (f->count_mode) {
            // counting mode: everything above is the EVALUATED part of the draw (it must exist, if only as a constant-0
            // record, to carry END_EVAL); the single-split trees follow in COUNTABLE records -- evaluated like any
            // DUAL record when per-point results are wanted, skipped and counted when only per-draw sums are.
            if (hdr.empty()) open_group((int64_t)s * f->m + pos, 0u);
            hdr.back() |= TK_FLAG_END_EVAL;
            pack_and_emit(countable, TK_FLAG_COUNTABLE, false);
        }
        hdr.back() |= TK_FLAG_END_DRAW;
    }
    // concatenate the per-draw plans
    f->draw_group_off.assign((size_t)f->S + 1, 0);
    for (int32_t s = 0; s < f->S; ++s)
        f->draw_group_off[(size_t)s + 1] = f->draw_group_off[(size_t)s] + (int32_t)plans[(size_t)s].hdr.size();
    const size_t n_groups_total = (size_t)f->draw_group_off[(size_t)f->S];
    std::vector<int32_t>& group_tree_off = P.group_tree_off;
    std::vector<uint32_t>&hdr = P.hdr, &split = P.split, &desc = P.desc;
    group_tree_off.assign(n_groups_total, 0);
    hdr.assign(n_groups_total, 0u);
    split.assign(n_groups_total, 0u);
    desc.assign(n_groups_total * (size_t)DW, 0u);
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < f->S; ++s) {
        const size_t g0 = (size_t)f->draw_group_off[(size_t)s];
        const DrawPlan& pl = plans[(size_t)s];
        std::copy(pl.group_tree_off.begin(), pl.group_tree_off.end(), group_tree_off.begin() + g0);
        std::copy(pl.hdr.begin(), pl.hdr.end(), hdr.begin() + g0);
        std::copy(pl.split.begin(), pl.split.end(), split.begin() + g0);
        std::copy(pl.desc.begin(), pl.desc.end(), desc.begin() + g0 * (size_t)DW);
    }
    f->n_groups = (int64_t)group_tree_off.size();
    group_tree_off.push_back((int32_t)n_trees);
    // group g covers trees [group_tree_off[g], group_tree_off[g+1]) -- groups never span draws
    pt.lap("upload: group planner");
}

// Walk form of the forest (what eval_walk_kernel, the WALK records and the table builder read): pre-order node
// records, per-tree node / leaf offsets, leaf values by leaf ordinal.  Host only.
void pack_walk_form(const bartint_forest* f, WalkForm& W) {
    PhaseTimer pt;
    const int64_t n_trees = (int64_t)f->S * f->m;
    const int64_t nn = f->n_nodes;
    StageVec<uint2>& nodes = W.nodes;
    StageVec<int32_t>&tree_off = W.tree_off, &leaf_off = W.leaf_off;
    StageVec<double>& leaf_mu = W.leaf_mu;
    nodes.resize((size_t)nn);
    tree_off.resize((size_t)n_trees + 1);
    leaf_off.assign((size_t)n_trees + 1, 0);
    leaf_mu.resize((size_t)f->n_leaves);
    for (int64_t k = 0; k <= n_trees; ++k) tree_off[(size_t)k] = (int32_t)f->tree_off[(size_t)k];
    // leaf offsets (prefix sum), then fill
    for (int64_t k = 0; k < n_trees; ++k) {
        int32_t cnt = tree_off[(size_t)k + 1] - tree_off[(size_t)k];
        leaf_off[(size_t)k + 1] = leaf_off[(size_t)k] + (cnt + 1) / 2;
    }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_trees; ++k) {
        const int32_t base = tree_off[(size_t)k], cnt = tree_off[(size_t)k + 1] - base;
        int32_t ord = 0;
        for (int32_t a = 0; a < cnt; ++a) {
            const size_t g = (size_t)(base + a);
            if (f->var[g] >= 0) {
                nodes[g] = make_uint2((uint32_t)f->var[g] | ((uint32_t)f->cut[g] << 16), (uint32_t)(f->right[g] - a));
            } else {
                nodes[g] = make_uint2(LEAF_VAR, (uint32_t)ord);
                leaf_mu[(size_t)leaf_off[(size_t)k] + ord] = f->mu[g];
                ++ord;
            }
        }
    }
    pt.lap("upload: pack walk form");
}

int forest_upload(bartint_forest* f, WalkForm* prebuilt) {
    PhaseTimer pt;
    WalkForm W_local;
    WalkForm& W = prebuilt ? *prebuilt : W_local;
    if (!prebuilt) pack_walk_form(f, W);
    const StageVec<uint2>& nodes = W.nodes;
    const StageVec<int32_t>&tree_off = W.tree_off, &leaf_off = W.leaf_off;
    const StageVec<double>& leaf_mu = W.leaf_mu;
    int rc;
    pt.lap("");                                      // restart: pack_walk_form reports its own time
    if ((rc = upload(&f->dev.nodes, nodes))) return rc;
    if ((rc = upload(&f->dev.tree_off, tree_off))) return rc;
    if ((rc = upload(&f->dev.tree_leaf_off, leaf_off))) return rc;
    if ((rc = upload(&f->dev.leaf_mu, leaf_mu))) return rc;
    if ((rc = upload(&f->dev.cut_off, f->cut_off))) return rc;
    if ((rc = upload(&f->dev.cut_val, f->cut_val))) return rc;

    if ((rc = launch_build_bitslice(f, nullptr))) return rc;
    pt.lap("upload: H2D walk form");
    FastPlan P;
    if (f->no_fast_plan || f->walk_only) f->table_ok = false;
    else plan_fast_path(f, tree_off.data(), P);
    pt.lap("");
    if (!f->table_ok) {
        if (f->defer_upload_sync) return BARTINT_OK;
        BI_CUDA(cudaStreamSynchronize(nullptr));
        return BARTINT_OK;
    }
    const int DW = P.DW;
    const std::vector<uint32_t>&desc = P.desc, &hdr = P.hdr, &split = P.split;
    if ((rc = upload(&f->dev.node_bit, P.node_bit))) return rc;
    if ((rc = upload(&f->dev.tree_perm, P.tree_perm))) return rc;
    if ((rc = upload(&f->dev.group_tree_off, P.group_tree_off))) return rc;
    if ((rc = upload(&f->dev.draw_group_off, f->draw_group_off))) return rc;
    BI_CUDA(cudaMallocAsync((void**)&f->dev.group_rec, (size_t)f->n_groups * (size_t)fast_rec_bytes(f) + 16, nullptr));
    uint32_t *d_desc = nullptr, *d_hdr = nullptr, *d_split = nullptr;
    if ((rc = upload(&d_desc, desc))) return rc;
    if ((rc = upload(&d_hdr, hdr))) { cudaFreeAsync(d_desc, nullptr); return rc; }
    if ((rc = upload(&d_split, split))) { cudaFreeAsync(d_desc, nullptr); cudaFreeAsync(d_hdr, nullptr); return rc; }
    rc = launch_build_tables(f, d_desc, DW, d_hdr, d_split, 0);
    cudaFreeAsync(d_desc, nullptr);
    cudaFreeAsync(d_hdr, nullptr);
    cudaFreeAsync(d_split, nullptr);
    // Surface copy / build errors here.  (Not needed for the host vectors: a cudaMemcpyAsync from pageable memory
    // returns once the source has been staged.)  The pipelined design step defers this wait: it orders its own
    // stream after the default stream with an event and lets the host run ahead to the next chunk of draws.
    cudaError_t e = f->defer_upload_sync ? cudaSuccess : cudaStreamSynchronize(nullptr);
    pt.lap("upload: H2D groups + build tables");
    if (rc) return rc;
    BI_CUDA(e);
    return BARTINT_OK;
}

}  // namespace bartint
