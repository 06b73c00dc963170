// C ABI of libbartint_cuda.so (see include/bartint.h for what each entry point replaces in the reference).
#include <cmath>
#include <cstring>
#include <ctime>
#include <cstdio>
#include <limits>
#include <memory>
#include <new>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.cuh"
#ifdef BARTINT_TEST_HOOKS
#include "../../include/bartint_test_hooks.h"
#endif

namespace bartint {
const char* get_error();

static int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s): libbartint_cuda has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return BARTINT_ERR_NODEVICE;
    }
    return BARTINT_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            active = true;
        }
    }
    ~DeviceGuard() {
        if (active) cudaSetDevice(prev);
    }
};

static void free_device(DeviceForest& d, unsigned char* image_base = nullptr) {
    for (void* q : {(void*)d.bs_pool, (void*)d.bs_slice_hdr, (void*)d.bs_base, (void*)d.bs_slot_draw})     // never part of an image
        if (q) cudaFreeAsync(q, nullptr);
    if (image_base != nullptr) {           // an unpacked forest: one allocation holds every other array
        cudaFreeAsync(image_base, nullptr);
        d = DeviceForest{};
        return;
    }
    void* ptrs[] = {d.nodes, d.tree_off, d.tree_leaf_off, d.leaf_mu, d.cut_off, d.cut_val, d.node_bit, d.tree_perm, d.group_tree_off,
                    d.draw_group_off, d.group_rec};
    for (void* q : ptrs)
        if (q) cudaFreeAsync(q, nullptr);      // ordered after every kernel already enqueued on the default stream
    d = DeviceForest{};
}

static int finish_forest(bartint_forest* f, bartint_forest** out, WalkForm* prebuilt = nullptr) {
    int rc = prebuilt ? BARTINT_OK : forest_finish_host(f);      // (a prebuilt walk form comes with its statistics)
    if (rc == BARTINT_OK) rc = check_device();
    if (rc == BARTINT_OK) {
        cudaError_t e = cudaGetDevice(&f->device);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    }
    if (rc == BARTINT_OK) {
        // keep freed blocks of the stream-ordered allocator cached across synchronisations (temporaries are
        // allocated per call with cudaMallocAsync; the default threshold 0 would hand them back to the driver)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, f->device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
        rc = forest_upload(f, prebuilt);
    }
    if (rc == BARTINT_OK) {
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaEventCreate(&f->ev0[k]);
            if (e == cudaSuccess) e = cudaEventCreate(&f->ev1[k]);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_last, cudaEventDisableTiming);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__);
    }
    if (rc != BARTINT_OK) {
        bartint_forest_destroy(f);
        return rc;
    }
    *out = f;
    return BARTINT_OK;
}

static int check_cuts(int32_t d, const int32_t* cut_offsets, const double* cut_values) {
    BI_REQUIRE(d >= 0, BARTINT_ERR_ARG, "d must be >= 0");
    BI_REQUIRE(cut_offsets != nullptr, BARTINT_ERR_ARG, "cut_offsets is NULL");
    BI_REQUIRE(cut_offsets[0] == 0, BARTINT_ERR_ARG, "cut_offsets[0] must be 0");
    for (int32_t j = 0; j < d; ++j)
        BI_REQUIRE(cut_offsets[j + 1] >= cut_offsets[j], BARTINT_ERR_ARG, "cut_offsets must be non-decreasing");
    BI_REQUIRE(cut_offsets[d] == 0 || cut_values != nullptr, BARTINT_ERR_ARG, "cut_values is NULL");
    return BARTINT_OK;
}

}  // namespace bartint

using namespace bartint;

extern "C" {

int bartint_version(void) { return 200; }

int64_t bartint_kernel_launch_count(void) { return (int64_t)bartint::launch_count(); }

const char* bartint_last_error(void) { return bartint::get_error(); }

int bartint_device_count(int* n_out) {
    BI_REQUIRE(n_out != nullptr, BARTINT_ERR_ARG, "n_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *n_out = n;
    return BARTINT_OK;
}

int bartint_set_host_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int bartint_set_device(int device) {
    int rc = check_device();
    if (rc) return rc;
    BI_CUDA(cudaSetDevice(device));
    return BARTINT_OK;
}

static int forest_from_dbarts_impl(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                   int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                   double y_min, double y_max, int32_t m, int32_t S, bool defer_upload_sync,
                                   bartint_forest** out, bool no_fast_plan = false);
static int forest_from_image(const void* host_header, int64_t host_bytes, const void* d_blob, int src_device,
                             int64_t device_bytes, cudaStream_t st, bool wait, bartint_forest** out);

int bartint_forest_from_dbarts098(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                  int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                  double y_min, double y_max, int32_t m, int32_t S, bartint_forest** out) {
    return forest_from_dbarts_impl(tree_strings, tree_fits, x_train, n, d, cut_offsets, cut_values, y_min, y_max, m, S,
                                   false, out);
}

static int forest_from_dbarts_impl(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                   int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                   double y_min, double y_max, int32_t m, int32_t S, bool defer_upload_sync,
                                   bartint_forest** out, bool no_fast_plan) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(S >= 0 && m >= 0 && n >= 0, BARTINT_ERR_ARG, "S, m, n must be >= 0");
    BI_REQUIRE((int64_t)S * m == 0 || tree_strings != nullptr, BARTINT_ERR_ARG, "tree_strings is NULL");
    BI_REQUIRE((int64_t)S * m * n == 0 || tree_fits != nullptr, BARTINT_ERR_ARG, "tree_fits is NULL");
    BI_REQUIRE(n * d == 0 || x_train != nullptr, BARTINT_ERR_ARG, "x_train is NULL");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest* f = new (std::nothrow) bartint_forest();
    BI_REQUIRE(f != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    f->S = S; f->m = m; f->d = d; f->y_min = y_min; f->y_max = y_max;
    f->defer_upload_sync = defer_upload_sync;
    f->no_fast_plan = no_fast_plan;
    f->cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f->cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    static const bool fused = [] { const char* e = std::getenv("BARTINT_FUSED_EXPORT"); return !(e && e[0] == '0'); }();
    if (no_fast_plan && fused) {      // the design step with few candidates: strings -> walk form in one pass
        WalkForm W;
        rc = parse_dbarts_walk(tree_strings, tree_fits, x_train, n, d, cut_offsets, cut_values, m, S, f, W);
        if (rc) { delete f; return rc; }
        return finish_forest(f, out, &W);
    }
    rc = parse_dbarts_forest(tree_strings, tree_fits, x_train, n, d, cut_offsets, cut_values, m, S, f);
    if (rc) { delete f; return rc; }
    return finish_forest(f, out);
}

int bartint_forest_from_soa(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset, const int32_t* split_var,
                            const int32_t* split_cut_idx, const int32_t* left, const int32_t* right,
                            const double* leaf_mu, const int32_t* cut_offsets, const double* cut_values, double y_min,
                            double y_max, bartint_forest** out) {
    return bartint_forest_from_soa_routed(S, m, d, tree_node_offset, split_var, split_cut_idx, left, right, leaf_mu,
                                          cut_offsets, cut_values, y_min, y_max, BARTINT_ROUTE_DBARTS, out);
}

int bartint_forest_from_wbart(const char* trees_text, int64_t text_len, int32_t d, const int32_t* cut_offsets,
                              const double* cut_values, double offset, bartint_forest** out) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(trees_text != nullptr, BARTINT_ERR_ARG, "trees_text is NULL");
    if (text_len <= 0) text_len = (int64_t)std::strlen(trees_text);
    BI_REQUIRE(trees_text[text_len] == '\0', BARTINT_ERR_ARG, "trees_text must be NUL terminated at text_len");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest* f = new (std::nothrow) bartint_forest();
    BI_REQUIRE(f != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    f->d = d; f->y_min = offset - 0.5; f->y_max = offset + 0.5; f->route = BARTINT_ROUTE_BART;
    f->cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f->cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    rc = parse_wbart_forest(trees_text, text_len, d, f);
    if (rc) { delete f; return rc; }
    return finish_forest(f, out);
}

int bartint_forest_route(const bartint_forest* f) { return f ? f->route : -1; }

int bartint_forest_from_value_splits(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                     const int32_t* split_var, const double* value, const int32_t* left,
                                     const int32_t* right, int32_t route, double leaf_scale, double offset,
                                     bartint_forest** out) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(route == BARTINT_ROUTE_DBARTS || route == BARTINT_ROUTE_BART, BARTINT_ERR_ARG, "unknown routing rule %d", route);
    BI_REQUIRE(S >= 0 && m >= 0 && d >= 0, BARTINT_ERR_ARG, "S, m and d must be >= 0");
    BI_REQUIRE(tree_node_offset != nullptr, BARTINT_ERR_ARG, "tree_node_offset is NULL");
    BI_REQUIRE((int64_t)S * m == 0 || (split_var && value), BARTINT_ERR_ARG, "node arrays must not be NULL");
    BI_REQUIRE((left == nullptr) == (right == nullptr), BARTINT_ERR_ARG, "left and right must both be given or both be NULL");
    bartint_forest* f = new (std::nothrow) bartint_forest();
    BI_REQUIRE(f != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    f->S = S; f->m = m; f->d = d; f->y_min = offset - 0.5; f->y_max = offset + 0.5; f->route = route;
    int rc = value_splits_to_soa(S, m, d, tree_node_offset, split_var, value, left, right, leaf_scale, f);
    if (rc) { delete f; return rc; }
    return finish_forest(f, out);
}

int bartint_forest_cutpoints(const bartint_forest* f, int32_t* cut_offsets, double* cut_values) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    if (cut_offsets) memcpy(cut_offsets, f->cut_off.data(), f->cut_off.size() * sizeof(int32_t));
    if (cut_values && !f->cut_val.empty()) memcpy(cut_values, f->cut_val.data(), f->cut_val.size() * sizeof(double));
    return BARTINT_OK;
}

int bartint_forest_from_soa_routed(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                   const int32_t* split_var, const int32_t* split_cut_idx, const int32_t* left,
                                   const int32_t* right, const double* leaf_mu, const int32_t* cut_offsets,
                                   const double* cut_values, double y_min, double y_max, int32_t route,
                                   bartint_forest** out) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(route == BARTINT_ROUTE_DBARTS || route == BARTINT_ROUTE_BART, BARTINT_ERR_ARG, "unknown routing rule %d", route);
    BI_REQUIRE(S >= 0 && m >= 0, BARTINT_ERR_ARG, "S and m must be >= 0");
    BI_REQUIRE(tree_node_offset != nullptr, BARTINT_ERR_ARG, "tree_node_offset is NULL");
    const bool has_nodes = (int64_t)S * m > 0;
    BI_REQUIRE(!has_nodes || (split_var && split_cut_idx && left && right && leaf_mu), BARTINT_ERR_ARG,
               "node arrays must not be NULL");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest* f = new (std::nothrow) bartint_forest();
    BI_REQUIRE(f != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    f->S = S; f->m = m; f->d = d; f->y_min = y_min; f->y_max = y_max; f->route = route;
    f->cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f->cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    rc = canonicalise_soa(S, m, d, tree_node_offset, split_var, split_cut_idx, left, right, leaf_mu, f);
    if (rc) { delete f; return rc; }
    return finish_forest(f, out);
}

#ifdef BARTINT_TEST_HOOKS      // include/bartint_test_hooks.h
int bartint_debug_plan_soa(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset, const int32_t* split_var,
                           const int32_t* split_cut_idx, const int32_t* left, const int32_t* right,
                           const double* leaf_mu, const int32_t* cut_offsets, const double* cut_values,
                           int64_t group_capacity, int64_t info[8], uint32_t* walk_nodes, int32_t* tree_leaf_offset,
                           double* leaf_mu_out, uint8_t* node_bit, int32_t* tree_perm, int32_t* draw_group_offset,
                           int32_t* group_tree_offset, uint32_t* hdr, uint32_t* split, uint32_t* desc) {
    BI_REQUIRE(info != nullptr, BARTINT_ERR_ARG, "info is NULL");
    BI_REQUIRE(S >= 0 && m >= 0, BARTINT_ERR_ARG, "S and m must be >= 0");
    BI_REQUIRE(tree_node_offset != nullptr, BARTINT_ERR_ARG, "tree_node_offset is NULL");
    const bool has_nodes = (int64_t)S * m > 0;
    BI_REQUIRE(!has_nodes || (split_var && split_cut_idx && left && right && leaf_mu), BARTINT_ERR_ARG,
               "node arrays must not be NULL");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest f;                       // never uploaded: no device state to release
    f.S = S; f.m = m; f.d = d;
    f.cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f.cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    if ((rc = canonicalise_soa(S, m, d, tree_node_offset, split_var, split_cut_idx, left, right, leaf_mu, &f))) return rc;
    if ((rc = forest_finish_host(&f))) return rc;
    WalkForm W;
    pack_walk_form(&f, W);
    FastPlan P;
    plan_fast_path(&f, W.tree_off.data(), P);
    info[0] = f.table_ok ? 1 : 0; info[1] = f.gather ? 1 : 0; info[2] = f.nb; info[3] = f.g_regs;
    info[4] = f.count_level; info[5] = f.n_groups; info[6] = P.DW; info[7] = f.n_leaves;
    BI_REQUIRE(f.n_groups <= group_capacity, BARTINT_ERR_ARG, "plan has %lld group records, capacity is %lld",
               (long long)f.n_groups, (long long)group_capacity);
    BI_REQUIRE(P.DW <= 16, BARTINT_ERR_UNSUPPORTED, "descriptor of %d words", P.DW);
    if (walk_nodes)
        for (size_t g = 0; g < W.nodes.size(); ++g) { walk_nodes[2 * g] = W.nodes[g].x; walk_nodes[2 * g + 1] = W.nodes[g].y; }
    auto copy_out = [](auto* dst, const auto& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    copy_out(tree_leaf_offset, W.leaf_off);
    copy_out(leaf_mu_out, W.leaf_mu);
    if (!f.table_ok) return BARTINT_OK;
    copy_out(node_bit, P.node_bit);
    copy_out(tree_perm, P.tree_perm);
    copy_out(draw_group_offset, f.draw_group_off);
    copy_out(group_tree_offset, P.group_tree_off);
    copy_out(hdr, P.hdr);
    copy_out(split, P.split);
    copy_out(desc, P.desc);
    return BARTINT_OK;
}

// shared tail of the two host-only export hooks: the host phases after parsing, then copy the SoA out
static int debug_finish_and_copy(bartint_forest& f, int64_t tree_capacity, int64_t node_capacity,
                                 int64_t* tree_node_offset, int32_t* split_var, int32_t* split_cut_idx, int32_t* left,
                                 int32_t* right, double* leaf_mu) {
    int rc = forest_finish_host(&f);
    if (rc) return rc;
    {
        WalkForm W;
        pack_walk_form(&f, W);
        FastPlan P;
        plan_fast_path(&f, W.tree_off.data(), P);
    }
    BI_REQUIRE(f.n_nodes <= node_capacity && (int64_t)f.S * f.m <= tree_capacity, BARTINT_ERR_ARG,
               "forest has %lld trees and %lld nodes, capacity is %lld and %lld", (long long)f.S * f.m,
               (long long)f.n_nodes, (long long)tree_capacity, (long long)node_capacity);
    const size_t nn = (size_t)f.n_nodes;
    if (tree_node_offset) memcpy(tree_node_offset, f.tree_off.data(), f.tree_off.size() * sizeof(int64_t));
    if (split_var && nn) memcpy(split_var, f.var.data(), nn * sizeof(int32_t));
    if (split_cut_idx && nn) memcpy(split_cut_idx, f.cut.data(), nn * sizeof(int32_t));
    if (left && nn) memcpy(left, f.left.data(), nn * sizeof(int32_t));
    if (right && nn) memcpy(right, f.right.data(), nn * sizeof(int32_t));
    if (leaf_mu && nn) memcpy(leaf_mu, f.mu.data(), nn * sizeof(double));
    return BARTINT_OK;
}

int bartint_debug_export_wbart(const char* trees_text, int64_t text_len, int32_t d, const int32_t* cut_offsets,
                               const double* cut_values, int64_t tree_capacity, int64_t node_capacity,
                               int64_t dims_out[3], int64_t* tree_node_offset, int32_t* split_var,
                               int32_t* split_cut_idx, int32_t* left, int32_t* right, double* leaf_mu) {
    BI_REQUIRE(dims_out != nullptr, BARTINT_ERR_ARG, "dims_out is NULL");
    dims_out[0] = dims_out[1] = dims_out[2] = 0;
    BI_REQUIRE(trees_text != nullptr, BARTINT_ERR_ARG, "trees_text is NULL");
    if (text_len <= 0) text_len = (int64_t)std::strlen(trees_text);
    BI_REQUIRE(trees_text[text_len] == '\0', BARTINT_ERR_ARG, "trees_text must be NUL terminated at text_len");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest f;                       // never uploaded: no device state to release
    f.d = d; f.route = BARTINT_ROUTE_BART;
    f.cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f.cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    if ((rc = parse_wbart_forest(trees_text, text_len, d, &f))) return rc;
    rc = debug_finish_and_copy(f, tree_capacity, node_capacity, tree_node_offset, split_var, split_cut_idx, left,
                               right, leaf_mu);
    dims_out[0] = f.S; dims_out[1] = f.m; dims_out[2] = f.n_nodes;
    return rc;
}

int bartint_debug_export_dbarts098_walk(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                        int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                        int32_t m, int32_t S, int64_t node_capacity, int64_t stats[8], uint32_t* walk_nodes,
                                        int32_t* tree_node_offset, int32_t* tree_leaf_offset, double* leaf_mu) {
    BI_REQUIRE(stats != nullptr, BARTINT_ERR_ARG, "stats is NULL");
    for (int k = 0; k < 8; ++k) stats[k] = 0;
    BI_REQUIRE(S >= 0 && m >= 0 && n >= 0, BARTINT_ERR_ARG, "S, m, n must be >= 0");
    BI_REQUIRE((int64_t)S * m == 0 || tree_strings != nullptr, BARTINT_ERR_ARG, "tree_strings is NULL");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest f;                       // never uploaded: no device state to release
    f.S = S; f.m = m; f.d = d;
    f.cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f.cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    WalkForm W;
    if ((rc = parse_dbarts_walk(tree_strings, tree_fits, x_train, n, d, cut_offsets, cut_values, m, S, &f, W))) return rc;
    stats[0] = f.n_nodes; stats[1] = f.n_leaves; stats[2] = f.max_depth; stats[3] = f.max_internal;
    stats[4] = f.draw_internal_max; stats[5] = f.bs_refs_total; stats[6] = f.bs_ok ? 1 : 0; stats[7] = f.bs_class_nodes[0];
    BI_REQUIRE(f.n_nodes <= node_capacity, BARTINT_ERR_ARG, "forest has %lld nodes, capacity is %lld", (long long)f.n_nodes,
               (long long)node_capacity);
    if (walk_nodes)
        for (size_t g = 0; g < W.nodes.size(); ++g) { walk_nodes[2 * g] = W.nodes[g].x; walk_nodes[2 * g + 1] = W.nodes[g].y; }
    if (tree_node_offset) memcpy(tree_node_offset, W.tree_off.data(), W.tree_off.size() * sizeof(int32_t));
    if (tree_leaf_offset) memcpy(tree_leaf_offset, W.leaf_off.data(), W.leaf_off.size() * sizeof(int32_t));
    if (leaf_mu && !W.leaf_mu.empty()) memcpy(leaf_mu, W.leaf_mu.data(), W.leaf_mu.size() * sizeof(double));
    return BARTINT_OK;
}

int bartint_debug_export_value_splits(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                      const int32_t* split_var, const double* value, const int32_t* left,
                                      const int32_t* right, double leaf_scale, int64_t* tree_node_offset_out,
                                      int32_t* var_out, int32_t* cut_out, int32_t* left_out, int32_t* right_out,
                                      double* leaf_mu_out, int32_t* cut_offsets_out, double* cut_values_out) {
    BI_REQUIRE(tree_node_offset != nullptr && (int64_t)S * m >= 0, BARTINT_ERR_ARG, "bad arguments");
    BI_REQUIRE((left == nullptr) == (right == nullptr), BARTINT_ERR_ARG, "left and right must both be given or both be NULL");
    bartint_forest f;                       // never uploaded: no device state to release
    f.S = S; f.m = m; f.d = d;
    int rc = value_splits_to_soa(S, m, d, tree_node_offset, split_var, value, left, right, leaf_scale, &f);
    if (rc) return rc;
    if (cut_offsets_out) memcpy(cut_offsets_out, f.cut_off.data(), f.cut_off.size() * sizeof(int32_t));
    if (cut_values_out && !f.cut_val.empty()) memcpy(cut_values_out, f.cut_val.data(), f.cut_val.size() * sizeof(double));
    return debug_finish_and_copy(f, (int64_t)S * m, tree_node_offset[(int64_t)S * m], tree_node_offset_out, var_out, cut_out,
                                 left_out, right_out, leaf_mu_out);
}

int bartint_debug_export_dbarts098(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                   int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                   int32_t m, int32_t S, int64_t node_capacity, int64_t* n_nodes_out,
                                   int64_t* tree_node_offset, int32_t* split_var, int32_t* split_cut_idx,
                                   int32_t* left, int32_t* right, double* leaf_mu) {
    BI_REQUIRE(n_nodes_out != nullptr, BARTINT_ERR_ARG, "n_nodes_out is NULL");
    *n_nodes_out = 0;
    BI_REQUIRE(S >= 0 && m >= 0 && n >= 0, BARTINT_ERR_ARG, "S, m, n must be >= 0");
    BI_REQUIRE((int64_t)S * m == 0 || tree_strings != nullptr, BARTINT_ERR_ARG, "tree_strings is NULL");
    BI_REQUIRE((int64_t)S * m * n == 0 || tree_fits != nullptr, BARTINT_ERR_ARG, "tree_fits is NULL");
    BI_REQUIRE(n * d == 0 || x_train != nullptr, BARTINT_ERR_ARG, "x_train is NULL");
    int rc = check_cuts(d, cut_offsets, cut_values);
    if (rc) return rc;
    bartint_forest f;                       // never uploaded: no device state to release
    f.S = S; f.m = m; f.d = d;
    f.cut_off.assign(cut_offsets, cut_offsets + d + 1);
    f.cut_val.assign(cut_values, cut_values + cut_offsets[d]);
    if ((rc = parse_dbarts_forest(tree_strings, tree_fits, x_train, n, d, cut_offsets, cut_values, m, S, &f))) return rc;
    rc = debug_finish_and_copy(f, (int64_t)S * m, node_capacity, tree_node_offset, split_var, split_cut_idx, left,
                               right, leaf_mu);
    *n_nodes_out = f.n_nodes;
    return rc;
}

#endif  // BARTINT_TEST_HOOKS

int bartint_forest_info(const bartint_forest* f, int64_t info[12]) {
    BI_REQUIRE(f != nullptr && info != nullptr, BARTINT_ERR_ARG, "NULL argument");
    info[0] = f->S; info[1] = f->m; info[2] = f->d; info[3] = f->n_nodes; info[4] = f->n_leaves;
    info[5] = f->max_depth; info[6] = f->max_internal; info[7] = f->table_ok ? 1 : 0; info[8] = f->n_groups;
    info[9] = f->nb;
    info[10] = f->table_ok ? (f->gather ? 2 : 1) : 0;
    info[11] = f->gather ? f->g_regs : 0;
    return BARTINT_OK;
}

int bartint_forest_bitslice_info(const bartint_forest* f, int64_t info[16]) {
    BI_REQUIRE(f != nullptr && info != nullptr, BARTINT_ERR_ARG, "NULL argument");
    for (int k = 0; k < 16; ++k) info[k] = 0;
    if (!f->bs_ok) return BARTINT_OK;
    info[0] = 1; info[1] = f->bs_ncuts; info[2] = f->bs_ch; info[3] = f->bs_rows; info[4] = f->bs_anc; info[5] = f->bs_q;
    info[6] = bs_pool_bytes(f); info[7] = (int64_t)bs_smem_bytes(f);
    for (int k = 0; k < 5; ++k) info[8 + k] = f->bs_class_nodes[k];
    info[13] = f->bs_refs_total;
    info[15] = (f->S + 31) / 32;
    { DeviceGuard g(f->device); info[14] = bs_padded_refs(f); }      // what the slices load: every class padded per slice
    return BARTINT_OK;
}

int bartint_forest_export_soa(const bartint_forest* f, int64_t* tree_node_offset, int32_t* split_var,
                              int32_t* split_cut_idx, int32_t* left, int32_t* right, double* leaf_mu) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    BI_REQUIRE(f->image_base == nullptr && !f->walk_only, BARTINT_ERR_UNSUPPORTED,
               "a forest restored from a device image keeps no host copy of the trees");
    const size_t nn = (size_t)f->n_nodes;
    if (tree_node_offset) memcpy(tree_node_offset, f->tree_off.data(), f->tree_off.size() * sizeof(int64_t));
    if (split_var && nn) memcpy(split_var, f->var.data(), nn * sizeof(int32_t));
    if (split_cut_idx && nn) memcpy(split_cut_idx, f->cut.data(), nn * sizeof(int32_t));
    if (left && nn) memcpy(left, f->left.data(), nn * sizeof(int32_t));
    if (right && nn) memcpy(right, f->right.data(), nn * sizeof(int32_t));
    if (leaf_mu && nn) memcpy(leaf_mu, f->mu.data(), nn * sizeof(double));
    return BARTINT_OK;
}

void bartint_forest_destroy(bartint_forest* f) {
    if (f == nullptr) return;
    for (bartint_forest* r : f->replicas)          // copies on the other members of the device group
        if (r && r != f) bartint_forest_destroy(r);
    f->replicas.clear();
    {
        DeviceGuard g(f->device);
        // the buffers are freed in default-stream order: make that stream wait for evaluations on other streams
        if (f->ev_last && f->ev_last_set) cudaStreamWaitEvent(nullptr, f->ev_last, 0);
        free_device(f->dev, f->image_base);
        f->image_base = nullptr;
        for (int k = 0; k < 2; ++k) {
            if (f->ev0[k]) cudaEventDestroy(f->ev0[k]);
            if (f->ev1[k]) cudaEventDestroy(f->ev1[k]);
        }
        if (f->ev_last) cudaEventDestroy(f->ev_last);
        (void)cudaGetLastError();
    }
    delete f;
}

// ------------------------------------------------------------------------------------------------------------
static int make_points(const bartint_forest* f, int64_t N, int32_t d, cudaStream_t st, bartint_points** out) {
    BI_REQUIRE(f != nullptr && out != nullptr, BARTINT_ERR_ARG, "NULL argument");
    *out = nullptr;
    BI_REQUIRE(N >= 0, BARTINT_ERR_ARG, "N must be >= 0");
    BI_REQUIRE(d == f->d, BARTINT_ERR_ARG, "point matrix has %d columns but the forest was fit on %d variables", d, f->d);
    bartint_points* p = new (std::nothrow) bartint_points();
    BI_REQUIRE(p != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    p->device = f->device; p->N = N; p->d = d;
    p->Npad = ((N + TK_NPAD - 1) / TK_NPAD) * TK_NPAD;
    if (p->Npad == 0) p->Npad = TK_NPAD;
    p->stream = st;
    DeviceGuard g(f->device);
    cudaError_t e = cudaMallocAsync((void**)&p->codes, (size_t)std::max(d, 1) * (size_t)p->Npad, st);
    if (e == cudaSuccess && f->gather)      // the gather kernel reads point-major packed codes (16 B per point)
        e = cudaMallocAsync((void**)&p->codes_pm, sizeof(uint4) * (size_t)p->Npad, st);
    if (e != cudaSuccess) {
        if (p->codes) cudaFreeAsync(p->codes, st);
        delete p;
        return cuda_fail(e, "cudaMallocAsync(codes)", __FILE__, __LINE__);
    }
    *out = p;
    return BARTINT_OK;
}

// The one-pass K0 of gather forests fills only the packed records; the feature-major rows are produced the first time
// a consumer needs them (walk kernel, table kernel, leaf decoding, bartint_points_codes), on that consumer's stream.
static int ensure_codes(const bartint_points* p, cudaStream_t st) {
    if (p->codes_valid) return BARTINT_OK;
    BI_REQUIRE(p->codes != nullptr && p->codes_pm != nullptr, BARTINT_ERR_ARG, "internal: point set has no codes");
    int rc = launch_unpack_codes(p->codes_pm, p->Npad, p->d, p->codes, st);
    if (rc == BARTINT_OK) p->codes_valid = true;
    return rc;
}

int bartint_points_from_device(const bartint_forest* f, const double* dX, int64_t N, int32_t d, int64_t ld,
                               void* stream, bartint_points** out) {
    BI_REQUIRE(N * d == 0 || dX != nullptr, BARTINT_ERR_ARG, "X is NULL");
    BI_REQUIRE(ld >= N, BARTINT_ERR_ARG, "leading dimension smaller than N");
    int rc = make_points(f, N, d, (cudaStream_t)stream, out);
    if (rc) return rc;
    DeviceGuard g(f->device);
    static const bool split_k0 = [] { const char* e = std::getenv("BARTINT_K0"); return e && std::strcmp(e, "split") == 0; }();
    if ((*out)->codes_pm && d >= 1 && !split_k0) {
        // gather forests: one pass, fp64 matrix -> packed records; feature-major rows only on demand (ensure_codes).
        // (Two earlier one-pass kernels that wrote BOTH layouts lost to bin + repack: while the cut-table probes
        //  bounded the binning, and because 4 points x all variables per thread made a few waves of long CTAs.)
        rc = launch_bin_pack_points(dX, ld, N, d, f->dev.cut_off, f->dev.cut_val, f->max_cuts, f->route, (*out)->codes_pm,
                                    (*out)->Npad, (cudaStream_t)stream);
        (*out)->codes_valid = false;
    } else {                                          // table forests, or BARTINT_K0=split (A/B timing): bin, then repack
        rc = launch_bin_points(dX, ld, N, d, f->dev.cut_off, f->dev.cut_val, f->max_cuts, f->route, (*out)->codes,
                               (*out)->Npad, (cudaStream_t)stream);
        if (rc == BARTINT_OK && (*out)->codes_pm)
            rc = launch_repack_codes((*out)->codes, (*out)->Npad, d, (*out)->codes_pm, (cudaStream_t)stream);
    }
    if (rc) { bartint_points_destroy(*out); *out = nullptr; }
    return rc;
}

// BARTINT_TIMING=2: wall-clock stamps of GPU progress (host callbacks on the streams; CLOCK_MONOTONIC ms modulo 10^5, the
// clock Python's time.monotonic() reads), for the timeline of a design step across the members of a device group.
namespace {
struct Stamp { int k; int b; const char* what; };
bool stamps_on() { static const bool on = [] { const char* e = std::getenv("BARTINT_TIMING"); return e && e[0] == '2'; }(); return on; }
void CUDART_CB stamp_cb(void* p) {
    Stamp* s = static_cast<Stamp*>(p);
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double t = std::fmod(1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec, 1e5);
    fprintf(stderr, "[bartint-stamp] %.3f device %d: %s %d\n", t, s->k, s->what, s->b);
    delete s;
}
void stamp(cudaStream_t st, int k, const char* what, int b = 0) {
    if (stamps_on()) cudaLaunchHostFunc(st, stamp_cb, new Stamp{k, b, what});
}
}  // namespace

static cudaError_t staged_submit_block(bartint_staged* s, int b) {
    const int64_t r0 = (int64_t)b * s->block_rows, rows = std::min(s->block_rows, s->N - r0);
    cudaError_t e = cudaSuccess;
    if (rows > 0)          // (one pitched 2-D copy; d separate 1-D copies of the column pieces run at the same 55 GB/s)
        e = cudaMemcpy2DAsync(s->dX + r0, sizeof(double) * (size_t)s->N, s->host_X + r0, sizeof(double) * (size_t)s->host_ld,
                              sizeof(double) * (size_t)rows, (size_t)s->d, cudaMemcpyHostToDevice, s->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(s->block_done[b], s->copy_stream);
    stamp(s->copy_stream, s->device, "X block landed", b);
    return e;
}

static cudaError_t staged_submit_rest(const bartint_staged* cs) {
    bartint_staged* s = const_cast<bartint_staged*>(cs);      // the handle is logically const for its consumers
    cudaError_t e = cudaSuccess;
    if (s->submitted >= s->n_blocks) return e;
    for (int b = s->submitted; b < s->n_blocks && e == cudaSuccess; ++b) e = staged_submit_block(s, b);
    s->submitted = s->n_blocks;
    if (e == cudaSuccess) e = cudaEventRecord(s->done, s->copy_stream);
    return e;
}

// The copy stream and the events of a staged matrix are recycled per host thread and device: creating a stream and five
// events costs ~0.1 ms per matrix, twice per sequential-design step.
namespace {
struct StageRes { int device; cudaStream_t st; cudaEvent_t done; cudaEvent_t blk[8]; };
thread_local std::vector<StageRes> g_stage_res;
}  // namespace

// rows [0, N) of a column-major host matrix with leading dimension ld, staged on `device`
static int stage_rows(const double* X, int64_t N, int64_t ld, int32_t d, int device, bartint_staged** out) {
    DeviceGuard g(device);
    bartint_staged* s = new (std::nothrow) bartint_staged();
    BI_REQUIRE(s != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    s->device = device; s->N = N; s->d = d; s->host_ld = ld;
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < g_stage_res.size(); ++k)
        if (g_stage_res[k].device == device) {                    // a recycled stream + events of this thread and device
            s->copy_stream = g_stage_res[k].st; s->done = g_stage_res[k].done;
            for (int b = 0; b < 8; ++b) s->block_done[b] = g_stage_res[k].blk[b];
            g_stage_res.erase(g_stage_res.begin() + (std::ptrdiff_t)k);
            break;
        }
    if (s->copy_stream == nullptr) {
        e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&s->dX, sizeof(double) * (size_t)std::max<int64_t>(N * d, 1), s->copy_stream);
    if (e == cudaSuccess && N * d > 0) {
        if (N >= (int64_t)1 << 18) {      // 4 row blocks (8 from 160 MB): column-major, so each block is a d-row 2-D copy
            // (the evaluation of the LAST block is the tail no transfer hides, and the first forest upload queues behind
            //  the block in flight on the copy engine: 8*10^6 x 10 rows in four 160 MB blocks had 0.9 ms and 3 ms of those)
            const int nb = N * (int64_t)d * 8 >= ((int64_t)160 << 20) ? 8 : 4;
            s->n_blocks = nb;
            s->block_rows = ((N + nb - 1) / nb + TK_NPAD - 1) / TK_NPAD * TK_NPAD;
            s->host_X = X;
            for (int b = 0; b < s->n_blocks && e == cudaSuccess; ++b)
                if (s->block_done[b] == nullptr) e = cudaEventCreateWithFlags(&s->block_done[b], cudaEventDisableTiming);
            // two of the four blocks travel at once (about as long as the export of the first draw chunk takes); the
            // rest follows the first forest upload, which must not queue behind the whole matrix on the copy engine
            static const int upfront = [] { const char* e2 = std::getenv("BARTINT_STAGE_UPFRONT"); int v = e2 ? atoi(e2) : 2; return v < 1 ? 1 : v > 4 ? 4 : v; }();
            for (int b = 0; b < upfront && e == cudaSuccess; ++b) e = staged_submit_block(s, b);
            s->submitted = upfront;
        } else if (ld == N) {
            e = cudaMemcpyAsync(s->dX, X, sizeof(double) * (size_t)(N * d), cudaMemcpyHostToDevice, s->copy_stream);
        } else {
            e = cudaMemcpy2DAsync(s->dX, sizeof(double) * (size_t)N, X, sizeof(double) * (size_t)ld, sizeof(double) * (size_t)N,
                                  (size_t)d, cudaMemcpyHostToDevice, s->copy_stream);
        }
    }
    if (e == cudaSuccess) e = cudaEventRecord(s->done, s->copy_stream);
    if (e != cudaSuccess) { bartint_staged_free(s); return cuda_fail(e, "bartint_stage_matrix", __FILE__, __LINE__); }
    *out = s;
    return BARTINT_OK;
}

int bartint_stage_matrix(const double* X, int64_t N, int32_t d, int device, bartint_staged** out) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(N >= 0 && d >= 0 && (N * d == 0 || X != nullptr), BARTINT_ERR_ARG, "X is NULL or bad shape");
    int rc = check_device();
    if (rc) return rc;
    if (device != BARTINT_DEVICE_GROUP) return stage_rows(X, N, N, d, device, out);
    // the device group: contiguous row shards (multiples of 1024 rows), shard g staged on member g by its own thread
    const int G = group_size();
    bartint_staged* box = new (std::nothrow) bartint_staged();
    BI_REQUIRE(box != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    box->device = group_device(0); box->N = N; box->d = d; box->n_blocks = 0;
    box->shards.assign((size_t)G, nullptr);
    box->shard_lo.assign((size_t)G + 1, N);
    const int64_t per = ((N + G - 1) / G + 1023) / 1024 * 1024;
    for (int g = 0; g <= G; ++g) box->shard_lo[(size_t)g] = std::min<int64_t>(N, (int64_t)g * per);
    for (int g = 1; g < G; ++g) {
        const int64_t lo = box->shard_lo[(size_t)g], rows = box->shard_lo[(size_t)g + 1] - lo;
        bartint_staged** slot = &box->shards[(size_t)g];
        // (the members other than 0 receive their forests over NVLink, not over the bus: nothing has to overtake the
        //  matrix on their copy engines, so the whole shard is put on the bus at once)
        group_post(g, [=] {
            int r = stage_rows(X + lo, rows, N, d, group_device(g), slot);
            if (r == BARTINT_OK && *slot != nullptr && staged_submit_rest(*slot) != cudaSuccess) r = BARTINT_ERR_CUDA;
            return r;
        });
    }
    rc = stage_rows(X, box->shard_lo[1], N, d, group_device(0), &box->shards[0]);
    const int rc2 = group_wait();
    if (rc == BARTINT_OK) rc = rc2;
    if (rc != BARTINT_OK) { bartint_staged_free(box); return rc; }
    *out = box;
    return BARTINT_OK;
}

int bartint_staged_flush(bartint_staged* s) {
    BI_REQUIRE(s != nullptr, BARTINT_ERR_ARG, "NULL argument");
    if (!s->shards.empty()) {
        for (bartint_staged* sh : s->shards) {
            int rc = bartint_staged_flush(sh);
            if (rc) return rc;
        }
        return BARTINT_OK;
    }
    DeviceGuard g(s->device);
    BI_CUDA(staged_submit_rest(s));
    return BARTINT_OK;
}

int bartint_staged_points(const bartint_forest* f, const bartint_staged* s, void* stream, bartint_points** out) {
    BI_REQUIRE(f != nullptr && s != nullptr && out != nullptr, BARTINT_ERR_ARG, "NULL argument");
    BI_REQUIRE(s->shards.empty(), BARTINT_ERR_UNSUPPORTED, "a matrix staged on the device group is consumed by bartint_design_step_dbarts098 / bartint_eval_staged");
    BI_REQUIRE(s->device == f->device, BARTINT_ERR_ARG, "matrix was staged on another device than the forest's");
    DeviceGuard g(f->device);
    BI_CUDA(staged_submit_rest(s));
    BI_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, s->done, 0));
    return bartint_points_from_device(f, s->dX, s->N, s->d, s->N, stream, out);
}

void bartint_staged_free(bartint_staged* s) {
    if (s == nullptr) return;
    if (!s->shards.empty()) {              // a group container: every shard is released by its member's thread
        for (size_t g = 1; g < s->shards.size(); ++g) {
            bartint_staged* sh = s->shards[g];
            if (sh && (int)g < group_size()) group_post((int)g, [sh] { bartint_staged_free(sh); return (int)BARTINT_OK; });
            else bartint_staged_free(sh);
        }
        bartint_staged_free(s->shards[0]);
        group_wait();
        delete s;
        return;
    }
    {
        DeviceGuard g(s->device);
        if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
        if (s->dX) cudaFree(s->dX);            // synchronous: a consumer stream may still be binning from it
        if (s->copy_stream && s->done && g_stage_res.size() < 16) {      // keep the stream and the events for the next matrix
            StageRes r{s->device, s->copy_stream, s->done, {}};
            for (int b = 0; b < 8; ++b) r.blk[b] = s->block_done[b];
            g_stage_res.push_back(r);
        } else {
            if (s->done) cudaEventDestroy(s->done);
            for (cudaEvent_t ev : s->block_done)
                if (ev) cudaEventDestroy(ev);
            if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
        }
        (void)cudaGetLastError();
    }
    delete s;
}

int bartint_points_from_host(const bartint_forest* f, const double* X, int64_t N, int32_t d, bartint_points** out) {
    BI_REQUIRE(f != nullptr && out != nullptr, BARTINT_ERR_ARG, "NULL argument");
    *out = nullptr;
    BI_REQUIRE(N >= 0 && (N * d == 0 || X != nullptr), BARTINT_ERR_ARG, "X is NULL or N < 0");
    DeviceGuard g(f->device);
    double* dX = nullptr;
    const size_t bytes = (size_t)std::max<int64_t>(N * d, 1) * sizeof(double);
    BI_CUDA(cudaMallocAsync((void**)&dX, bytes, nullptr));
    cudaError_t e = (N * d) ? cudaMemcpy(dX, X, (size_t)(N * d) * sizeof(double), cudaMemcpyHostToDevice) : cudaSuccess;
    int rc = (e == cudaSuccess) ? bartint_points_from_device(f, dX, N, d, N, nullptr, out)
                                : cuda_fail(e, "cudaMemcpy(X)", __FILE__, __LINE__);
    if (rc == BARTINT_OK) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { rc = cuda_fail(e, "bin_points", __FILE__, __LINE__); bartint_points_destroy(*out); *out = nullptr; }
    }
    cudaFreeAsync(dX, nullptr);
    return rc;
}

int bartint_points_generate(const bartint_forest* f, int32_t measure, int64_t N, uint64_t seed, int64_t first_index,
                            double* dX_out, void* stream, bartint_points** out) {
    BI_REQUIRE(f != nullptr && out != nullptr, BARTINT_ERR_ARG, "NULL argument");
    *out = nullptr;
    BI_REQUIRE(measure >= 0 && measure <= 3, BARTINT_ERR_ARG, "unknown measure %d", measure);
    BI_REQUIRE(N >= 0 && first_index >= 0, BARTINT_ERR_ARG, "N and first_index must be >= 0");
    int rc = make_points(f, N, f->d, (cudaStream_t)stream, out);
    if (rc) return rc;
    DeviceGuard g(f->device);
    rc = launch_generate_points(measure, N, first_index, seed, f->d, f->dev.cut_off, f->dev.cut_val, f->route,
                                (*out)->codes, (*out)->Npad, dX_out, (cudaStream_t)stream);
    if (rc == BARTINT_OK && (*out)->codes_pm)
        rc = launch_repack_codes((*out)->codes, (*out)->Npad, f->d, (*out)->codes_pm, (cudaStream_t)stream);
    if (rc) { bartint_points_destroy(*out); *out = nullptr; }
    return rc;
}

int64_t bartint_points_count(const bartint_points* p) { return p ? p->N : -1; }

int bartint_points_codes(const bartint_points* p, uint8_t* codes_out) {
    BI_REQUIRE(p != nullptr && (codes_out != nullptr || p->N * p->d == 0), BARTINT_ERR_ARG, "NULL argument");
    DeviceGuard g(p->device);
    int rc = ensure_codes(p, nullptr);
    if (rc) return rc;
    BI_CUDA(cudaDeviceSynchronize());
    if (p->N > 0 && p->d > 0)
        BI_CUDA(cudaMemcpy2D(codes_out, (size_t)p->N, p->codes, (size_t)p->Npad, (size_t)p->N, (size_t)p->d,
                             cudaMemcpyDeviceToHost));
    return BARTINT_OK;
}

void bartint_points_destroy(bartint_points* p) {
    if (p == nullptr) return;
    {
        DeviceGuard g(p->device);
        if (p->codes) cudaFreeAsync(p->codes, p->stream);
        if (p->codes_pm) cudaFreeAsync(p->codes_pm, p->stream);
        (void)cudaGetLastError();
    }
    delete p;
}

// ------------------------------------------------------------------------------------------------------------
static int check_range(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end) {
    BI_REQUIRE(f != nullptr && p != nullptr, BARTINT_ERR_ARG, "NULL forest or points");
    BI_REQUIRE(p->d == f->d && p->device == f->device, BARTINT_ERR_ARG, "points were not created for this forest");
    BI_REQUIRE(0 <= draw_begin && draw_begin <= draw_end && draw_end <= f->S, BARTINT_ERR_ARG,
               "draw range [%d, %d) outside [0, %d)", draw_begin, draw_end, f->S);
    return BARTINT_OK;
}

// core: run the evaluation kernel + second-stage merges, everything on `st`, outputs on device (scaled units)
static int eval_core(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                     int32_t draw_end, double* d_draw_sum, double* d_point_mean, double* d_point_m2,
                     int32_t* d_leaf, double* d_full, double fp_scale, double fp_y0, int fp_scaled, cudaStream_t st) {
    const int ndraws = draw_end - draw_begin;
    const int slot = (d_point_mean != nullptr || d_full != nullptr || d_leaf != nullptr) ? 1 : 0;
    f->last_launches[slot] = 0;
    f->last_timed[slot] = false;
    if (ndraws == 0) return BARTINT_OK;
    const bool want_mom = d_point_mean != nullptr && d_point_m2 != nullptr;
    if (d_leaf != nullptr) flags |= BARTINT_FORCE_WALK;
    EvalPlan plan = make_plan(f, p, flags, draw_begin, draw_end, want_mom, d_full != nullptr);
    if (plan.use_bitslice ? p->codes_pm == nullptr : !plan.use_gather) {   // walk / table kernels read the feature-major rows
        int rc0 = ensure_codes(p, st);
        if (rc0) return rc0;
    }
    f->last_used_table[slot] = plan.use_bitslice ? 2 : plan.use_table ? 1 : 0;
    double *draw_part = nullptr, *pt_mean = nullptr, *pt_m2 = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&draw_part, sizeof(double) * (size_t)plan.n_tiles * (size_t)(plan.use_bitslice ? plan.chunk_draws : 1) * (size_t)ndraws, st));
    const bool direct_mom = want_mom && plan.n_chunks == 1;   // single chunk: the kernel writes the final moments
    if (want_mom) {
        if (direct_mom) { pt_mean = d_point_mean; pt_m2 = d_point_m2; }
        else {
            const size_t bytes = sizeof(double) * (size_t)plan.n_chunks * (size_t)std::max<int64_t>(p->N, 1);
            BI_CUDA(cudaMallocAsync((void**)&pt_mean, bytes, st));
            BI_CUDA(cudaMallocAsync((void**)&pt_m2, bytes, st));
        }
    }
    BI_CUDA(cudaEventRecord(f->ev0[slot], st));
    int rc = plan.use_bitslice ? launch_eval_bitslice(f, p, plan, draw_part, st)
             : plan.use_gather ? launch_eval_gather(f, p, plan, draw_part, pt_mean, pt_m2, d_full, fp_scale, fp_y0, fp_scaled, st)
             : plan.use_table ? launch_eval_table(f, p, plan, draw_part, pt_mean, pt_m2, d_full, fp_scale, fp_y0, fp_scaled, st)
                              : launch_eval_walk(f, p, plan, draw_part, pt_mean, pt_m2, d_leaf, d_full, fp_scale, fp_y0, fp_scaled, st);
    BI_CUDA(cudaEventRecord(f->ev1[slot], st));
    f->last_timed[slot] = true;
    f->last_launches[slot] = 1;
    if (rc == BARTINT_OK && d_draw_sum != nullptr && plan.use_bitslice) {
        rc = launch_bs_reduce(f, draw_part, plan.n_tiles * plan.chunk_draws, draw_begin, ndraws, p->N, d_draw_sum, st);
        f->last_launches[slot]++;
    } else if (rc == BARTINT_OK && d_draw_sum != nullptr) {
        rc = launch_reduce_draw_parts(draw_part, plan.n_tiles, ndraws, d_draw_sum, st);
        f->last_launches[slot]++;
        if (rc == BARTINT_OK && plan.count_mode) {      // + the single-split trees, from 1-D code histograms
            rc = launch_count_single_splits(f, p, draw_begin, draw_end, d_draw_sum, st);
            f->last_launches[slot] += 3;
        }
    }
    if (rc == BARTINT_OK && want_mom && !direct_mom) {
        rc = launch_merge_point_parts(pt_mean, pt_m2, plan.n_chunks, plan.chunk_draws, ndraws, p->N, d_point_mean,
                                      d_point_m2, st);
        f->last_launches[slot]++;
    }
    cudaFreeAsync(draw_part, st);
    if (want_mom && !direct_mom) { cudaFreeAsync(pt_mean, st); cudaFreeAsync(pt_m2, st); }
    if (cudaEventRecord(f->ev_last, st) == cudaSuccess) f->ev_last_set = true;
    return rc;
}

int bartint_eval_points_device(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                               int32_t draw_end, double* d_draw_sum, double* d_point_mean, double* d_point_m2,
                               void* stream) {
    int rc = check_range(f, p, draw_begin, draw_end);
    if (rc) return rc;
    BI_REQUIRE((d_point_mean == nullptr) == (d_point_m2 == nullptr), BARTINT_ERR_ARG,
               "d_point_mean and d_point_m2 must both be given or both be NULL");
    DeviceGuard g(f->device);
    return eval_core(f, p, flags, draw_begin, draw_end, d_draw_sum, d_point_mean, d_point_m2, nullptr, nullptr, 1.0, 0.0,
                     1, (cudaStream_t)stream);
}

int bartint_eval_points_full_device(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                                    int32_t draw_end, double* d_full_pred, void* stream) {
    int rc = check_range(f, p, draw_begin, draw_end);
    if (rc) return rc;
    BI_REQUIRE(d_full_pred != nullptr || (int64_t)(draw_end - draw_begin) * p->N == 0, BARTINT_ERR_ARG, "d_full_pred is NULL");
    DeviceGuard g(f->device);
    const int scaled = (flags & BARTINT_SCALED_UNITS) ? 1 : 0;
    return eval_core(f, p, flags, draw_begin, draw_end, nullptr, nullptr, nullptr, nullptr, d_full_pred, f->y_max - f->y_min,
                     f->y_min, scaled, (cudaStream_t)stream);
}

int bartint_merge_moments_device(const bartint_forest* f, uint32_t flags, int32_t G, const int32_t* shard_draws,
                                 const double* d_mean, const double* d_m2, int64_t N, const double* d_weights,
                                 int64_t weights_len, double* d_point_mean_out, double* d_point_var_out, void* stream) {
    BI_REQUIRE(f != nullptr && shard_draws != nullptr && G >= 1, BARTINT_ERR_ARG, "bad shard description");
    BI_REQUIRE(N == 0 || (d_mean != nullptr && d_m2 != nullptr), BARTINT_ERR_ARG, "NULL moments");
    BI_REQUIRE(d_weights == nullptr || weights_len == 1 || weights_len == N, BARTINT_ERR_ARG,
               "weights must have length 1 or N");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* d_counts = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&d_counts, sizeof(int32_t) * (size_t)G, st));
    BI_CUDA(cudaMemcpyAsync(d_counts, shard_draws, sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice, st));
    const int scaled = (flags & BARTINT_SCALED_UNITS) ? 1 : 0;
    int rc = launch_merge_shards(d_mean, d_m2, G, d_counts, N, f->y_max - f->y_min, f->y_min, scaled, d_weights,
                                 weights_len, d_point_mean_out, d_point_var_out, st);
    cudaFreeAsync(d_counts, st);
    return rc;
}

int bartint_sum_rows_device(const double* d_rows, int32_t n_rows, int32_t n, double* d_out, void* stream) {
    BI_REQUIRE(n_rows >= 1 && n >= 0 && (n == 0 || (d_rows != nullptr && d_out != nullptr)), BARTINT_ERR_ARG, "bad arguments");
    return launch_reduce_draw_parts(d_rows, n_rows, n, d_out, (cudaStream_t)stream);
}

int bartint_argmax_ties_device(const double* d_values, int64_t n, int64_t* d_out, void* stream) {
    BI_REQUIRE(d_out != nullptr && n >= 0 && (n == 0 || d_values != nullptr), BARTINT_ERR_ARG, "bad arguments");
    return launch_argmax_ties(d_values, n, d_out, (cudaStream_t)stream);
}

int bartint_topk_device(const double* d_values, int64_t n, int32_t k, int64_t* d_index_out, double* d_value_out, void* stream) {
    BI_REQUIRE(k >= 0 && n >= 0 && (n == 0 || d_values != nullptr) && (k == 0 || d_index_out != nullptr), BARTINT_ERR_ARG, "bad arguments");
    return launch_topk(d_values, n, k, d_index_out, d_value_out, (cudaStream_t)stream);
}

int bartint_topk(const double* values, int64_t n, int32_t k, int64_t* index_out, double* value_out) {
    BI_REQUIRE(k >= 0 && n >= 0 && (n == 0 || values != nullptr) && (k == 0 || index_out != nullptr), BARTINT_ERR_ARG, "bad arguments");
    int rc = check_device();
    if (rc) return rc;
    if (k == 0) return BARTINT_OK;
    double* d_v = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&d_v, sizeof(double) * (size_t)(std::max<int64_t>(n, 1) + k) + sizeof(int64_t) * (size_t)k, nullptr));
    double* d_val = d_v + std::max<int64_t>(n, 1);
    int64_t* d_idx = reinterpret_cast<int64_t*>(d_val + k);
    cudaError_t e = n ? cudaMemcpyAsync(d_v, values, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, nullptr) : cudaSuccess;
    rc = e == cudaSuccess ? launch_topk(d_v, n, k, d_idx, d_val, nullptr) : cuda_fail(e, "bartint_topk", __FILE__, __LINE__);
    if (rc == BARTINT_OK) {
        e = cudaMemcpy(index_out, d_idx, sizeof(int64_t) * (size_t)k, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && value_out) e = cudaMemcpy(value_out, d_val, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "bartint_topk", __FILE__, __LINE__);
    }
    cudaFreeAsync(d_v, nullptr);
    return rc;
}

int bartint_eval_points(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                        int32_t draw_end, const double* weights, int64_t weights_len, double* draw_mean,
                        double* point_mean, double* point_var, double* full_pred) {
    int rc = check_range(f, p, draw_begin, draw_end);
    if (rc) return rc;
    BI_REQUIRE(weights == nullptr || weights_len == 1 || weights_len == p->N, BARTINT_ERR_ARG,
               "weights must have length 1 or N");
    DeviceGuard g(f->device);
    const int ndraws = draw_end - draw_begin;
    const int64_t N = p->N;
    const int scaled = (flags & BARTINT_SCALED_UNITS) ? 1 : 0;
    const double scale = f->y_max - f->y_min, y0 = f->y_min;
    cudaStream_t st = nullptr;
    const bool want_mom = point_mean != nullptr || point_var != nullptr;
    double *d_sum = nullptr, *d_mean = nullptr, *d_m2 = nullptr, *d_full = nullptr, *d_out = nullptr, *d_w = nullptr,
           *d_omean = nullptr, *d_ovar = nullptr;
    int32_t* d_cnt = nullptr;
    rc = BARTINT_OK;
    auto cleanup = [&]() {
        cudaFreeAsync(d_sum, nullptr); cudaFreeAsync(d_mean, nullptr); cudaFree(d_m2); cudaFreeAsync(d_full, nullptr); cudaFreeAsync(d_out, nullptr); cudaFreeAsync(d_w, nullptr);
        cudaFreeAsync(d_omean, nullptr); cudaFreeAsync(d_ovar, nullptr); cudaFreeAsync(d_cnt, nullptr);
    };
#define EP_CUDA(call)                                                             \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #call, __FILE__, __LINE__); } \
    } while (0)
    const size_t nb_draw = sizeof(double) * (size_t)std::max(ndraws, 1), nb_pt = sizeof(double) * (size_t)std::max<int64_t>(N, 1);
    EP_CUDA(cudaMallocAsync((void**)&d_sum, nb_draw, nullptr));
    EP_CUDA(cudaMallocAsync((void**)&d_out, nb_draw, nullptr));
    if (want_mom) {
        EP_CUDA(cudaMallocAsync((void**)&d_mean, nb_pt, nullptr));
        EP_CUDA(cudaMallocAsync((void**)&d_m2, nb_pt, nullptr));
        EP_CUDA(cudaMallocAsync((void**)&d_omean, nb_pt, nullptr));
        EP_CUDA(cudaMallocAsync((void**)&d_ovar, nb_pt, nullptr));
        EP_CUDA(cudaMallocAsync((void**)&d_cnt, sizeof(int32_t), nullptr));
        if (weights != nullptr) {
            EP_CUDA(cudaMallocAsync((void**)&d_w, sizeof(double) * (size_t)weights_len, nullptr));
            EP_CUDA(cudaMemcpy(d_w, weights, sizeof(double) * (size_t)weights_len, cudaMemcpyHostToDevice));
        }
    }
    if (full_pred != nullptr && (int64_t)ndraws * N > 0)
        EP_CUDA(cudaMallocAsync((void**)&d_full, sizeof(double) * (size_t)ndraws * (size_t)N, nullptr));
    rc = eval_core(f, p, flags, draw_begin, draw_end, d_sum, want_mom ? d_mean : nullptr, want_mom ? d_m2 : nullptr,
                   nullptr, d_full, scale, y0, scaled, st);
    if (rc == BARTINT_OK && draw_mean != nullptr && ndraws > 0)
        rc = launch_finalize_draws(d_sum, ndraws, N, scale, y0, scaled, d_out, st);
    if (rc == BARTINT_OK && want_mom && N > 0) {
        if (ndraws == 0) {
            // no draws: colMeans / colVars of a 0-row matrix are NaN
            cleanup();
            const double nan = std::numeric_limits<double>::quiet_NaN();
            for (int64_t i = 0; i < N; ++i) { if (point_mean) point_mean[i] = nan; if (point_var) point_var[i] = nan; }
            return BARTINT_OK;
        }
        EP_CUDA(cudaMemcpy(d_cnt, &ndraws, sizeof(int32_t), cudaMemcpyHostToDevice));
        rc = launch_merge_shards(d_mean, d_m2, 1, d_cnt, N, scale, y0, scaled, d_w, weights_len, d_omean, d_ovar, st);
    }
    if (rc != BARTINT_OK) { cleanup(); return rc; }
    EP_CUDA(cudaDeviceSynchronize());
    if (draw_mean != nullptr && ndraws > 0) EP_CUDA(cudaMemcpy(draw_mean, d_out, sizeof(double) * (size_t)ndraws, cudaMemcpyDeviceToHost));
    if (point_mean != nullptr && N > 0) EP_CUDA(cudaMemcpy(point_mean, d_omean, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
    if (point_var != nullptr && N > 0) EP_CUDA(cudaMemcpy(point_var, d_ovar, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
    if (d_full != nullptr) EP_CUDA(cudaMemcpy(full_pred, d_full, sizeof(double) * (size_t)ndraws * (size_t)N, cudaMemcpyDeviceToHost));
    cleanup();
#undef EP_CUDA
    return BARTINT_OK;
}

// One shard of a host-buffer evaluation on ONE device (the forest's): rows [0, N) of a column-major matrix with
// leading dimension ld, processed as a pipeline over row blocks -- the host->device copy of block k+1 (copy stream)
// overlaps the binning + evaluation of block k (compute stream); per-block draw sums are added in block order,
// per-point outputs (and the S x N matrix, block by block) land at their offsets.  d_sum_keep != NULL: leave the scaled
// per-draw sums of these rows there (device, S doubles) for a cross-device exchange instead of finalising draw_mean.
// Blocks until its streams are done.
static int eval_shard(const bartint_forest* f, const double* X, int64_t N, int64_t ld, int32_t d, uint32_t flags,
                      const double* weights, int64_t weights_len, double* d_sum_keep, double* draw_mean, double* point_mean,
                      double* point_var, double* full_pred) {
    DeviceGuard g(f->device);
    const int S = f->S;
    const int scaled = (flags & BARTINT_SCALED_UNITS) ? 1 : 0;
    const double scale = f->y_max - f->y_min, y0 = f->y_min;
    const bool want_mom = point_mean != nullptr || point_var != nullptr;
    const bool want_full = full_pred != nullptr && S > 0;
    const int K = 4;                                                   // row blocks
    int64_t blk = (((N + K - 1) / K + TK_NPAD - 1) / TK_NPAD) * TK_NPAD;
    if (want_full) blk = std::min<int64_t>(blk, std::max<int64_t>(TK_NPAD, ((int64_t)1 << 27) / std::max(S, 1) / TK_NPAD * TK_NPAD));   // <= 1 GiB per block buffer
    const int nblk = (int)std::max<int64_t>((N + blk - 1) / blk, 1);
    cudaStream_t cs = nullptr, xs = nullptr;       // compute, copy
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    double *dX[2] = {nullptr, nullptr}, *dF[2] = {nullptr, nullptr}, *d_sums = nullptr, *d_sum = nullptr, *d_out = nullptr,
           *d_mean = nullptr, *d_m2 = nullptr, *d_omean = nullptr, *d_ovar = nullptr, *d_w = nullptr;
    int32_t* d_cnt = nullptr;
    bartint_points* pts[2] = {nullptr, nullptr};
    int rc = BARTINT_OK;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t err) { if (e == cudaSuccess && err != cudaSuccess) e = err; return e == cudaSuccess; };
    const size_t n1 = (size_t)std::max<int64_t>(N, 1), s1 = (size_t)std::max(S, 1);
    ok(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&xs, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        ok(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&consumed[k], cudaEventDisableTiming));
        ok(cudaMallocAsync((void**)&dX[k], sizeof(double) * (size_t)blk * (size_t)std::max(d, 1), cs));
        if (want_full) ok(cudaMallocAsync((void**)&dF[k], sizeof(double) * (size_t)blk * s1, cs));
    }
    ok(cudaMallocAsync((void**)&d_sums, sizeof(double) * (size_t)nblk * s1, cs));
    ok(cudaMallocAsync((void**)&d_sum, sizeof(double) * s1, cs));
    ok(cudaMallocAsync((void**)&d_out, sizeof(double) * s1, cs));
    if (want_mom) {
        ok(cudaMallocAsync((void**)&d_mean, sizeof(double) * n1, cs));
        ok(cudaMallocAsync((void**)&d_m2, sizeof(double) * n1, cs));
        ok(cudaMallocAsync((void**)&d_omean, sizeof(double) * n1, cs));
        ok(cudaMallocAsync((void**)&d_ovar, sizeof(double) * n1, cs));
        ok(cudaMallocAsync((void**)&d_cnt, sizeof(int32_t), cs));
        if (weights != nullptr) {
            ok(cudaMallocAsync((void**)&d_w, sizeof(double) * (size_t)weights_len, cs));
            ok(cudaMemcpyAsync(d_w, weights, sizeof(double) * (size_t)weights_len, cudaMemcpyHostToDevice, cs));
        }
        ok(cudaMemcpyAsync(d_cnt, &S, sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    }
    ok(cudaStreamSynchronize(cs));                 // buffers exist before the copy stream touches them
    for (int k = 0; k < nblk && e == cudaSuccess && rc == BARTINT_OK && N > 0; ++k) {
        const int b = k & 1;
        const int64_t r0 = (int64_t)k * blk, rows = std::min<int64_t>(blk, N - r0);
        if (k >= 2) ok(cudaStreamWaitEvent(xs, consumed[b], 0));       // block k-2 has been binned out of dX[b]
        ok(cudaMemcpy2DAsync(dX[b], sizeof(double) * (size_t)rows, X + r0, sizeof(double) * (size_t)ld,
                             sizeof(double) * (size_t)rows, (size_t)d, cudaMemcpyHostToDevice, xs));
        ok(cudaEventRecord(copied[b], xs));
        ok(cudaStreamWaitEvent(cs, copied[b], 0));
        if (pts[b]) { bartint_points_destroy(pts[b]); pts[b] = nullptr; }
        if (e != cudaSuccess) break;
        rc = bartint_points_from_device(f, dX[b], rows, d, rows, cs, &pts[b]);
        ok(cudaEventRecord(consumed[b], cs));
        if (rc == BARTINT_OK)
            rc = eval_core(f, pts[b], flags, 0, S, d_sums + (size_t)k * s1, want_mom ? d_mean + r0 : nullptr,
                           want_mom ? d_m2 + r0 : nullptr, nullptr, want_full ? dF[b] : nullptr, scale, y0, scaled, cs);
        if (rc == BARTINT_OK && want_full)          // the block's S x rows slab is contiguous in R's layout
            ok(cudaMemcpyAsync(full_pred + (size_t)S * (size_t)r0, dF[b], sizeof(double) * (size_t)S * (size_t)rows,
                               cudaMemcpyDeviceToHost, cs));
    }
    if (e == cudaSuccess && rc == BARTINT_OK && N == 0) ok(cudaMemsetAsync(d_sums, 0, sizeof(double) * s1, cs));
    if (e == cudaSuccess && rc == BARTINT_OK && S > 0)
        rc = launch_reduce_draw_parts(d_sums, N > 0 ? nblk : 1, S, d_sum_keep ? d_sum_keep : d_sum, cs);
    if (e == cudaSuccess && rc == BARTINT_OK && draw_mean != nullptr && d_sum_keep == nullptr && S > 0)
        rc = launch_finalize_draws(d_sum, S, N, scale, y0, scaled, d_out, cs);
    if (e == cudaSuccess && rc == BARTINT_OK && want_mom && N > 0)
        rc = launch_merge_shards(d_mean, d_m2, 1, d_cnt, N, scale, y0, scaled, d_w, weights_len, d_omean, d_ovar, cs);
    if (e == cudaSuccess && rc == BARTINT_OK) {
        if (draw_mean && d_sum_keep == nullptr && S > 0) ok(cudaMemcpyAsync(draw_mean, d_out, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost, cs));
        if (point_mean && N > 0) ok(cudaMemcpyAsync(point_mean, d_omean, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, cs));
        if (point_var && N > 0) ok(cudaMemcpyAsync(point_var, d_ovar, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, cs));
    }
    if (cs) ok(cudaStreamSynchronize(cs));
    if (xs) cudaStreamSynchronize(xs);
    for (int k = 0; k < 2; ++k) {
        if (pts[k]) bartint_points_destroy(pts[k]);
        if (dX[k]) cudaFreeAsync(dX[k], cs);
        if (dF[k]) cudaFreeAsync(dF[k], cs);
        if (copied[k]) cudaEventDestroy(copied[k]);
        if (consumed[k]) cudaEventDestroy(consumed[k]);
    }
    double* tmp[] = {d_sums, d_sum, d_out, d_mean, d_m2, d_omean, d_ovar, d_w};
    for (double* q : tmp)
        if (q) cudaFreeAsync(q, cs);
    if (d_cnt) cudaFreeAsync(d_cnt, cs);
    if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
    if (xs) cudaStreamDestroy(xs);
    if (rc == BARTINT_OK && e != cudaSuccess) rc = cuda_fail(e, "bartint_eval pipeline", __FILE__, __LINE__);
    return rc;
}

// Copies of a forest on the other members of the device group (built on first use, kept with the forest): the image
// is packed once on the forest's device and pulled by every member over NVLink.
static int ensure_replicas(const bartint_forest* cf) {
    bartint_forest* f = const_cast<bartint_forest*>(cf);          // the replica list is a cache (one caller at a time)
    const int G = group_size();
    if ((int)f->replicas.size() == G && f->replica_generation == group_generation()) return BARTINT_OK;
    for (bartint_forest* r : f->replicas)
        if (r && r != f) bartint_forest_destroy(r);
    f->replicas.assign((size_t)G, nullptr);
    f->replicas[0] = f;
    f->replica_generation = group_generation();
    if (G == 1) return BARTINT_OK;
    DeviceGuard g(f->device);
    int64_t hb = 0, db = 0;
    int rc = bartint_forest_image_size(f, &hb, &db);
    if (rc) return rc;
    auto hdr = std::make_shared<std::vector<unsigned char>>((size_t)hb);
    void* blob = nullptr;
    BI_CUDA(cudaMallocAsync(&blob, (size_t)db, nullptr));
    rc = bartint_forest_pack(f, hdr->data(), hb, blob, db, nullptr);
    cudaEvent_t pev = nullptr;
    if (rc == BARTINT_OK && (cudaEventCreateWithFlags(&pev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(pev, nullptr) != cudaSuccess))
        rc = cuda_fail(cudaGetLastError(), "replicate forest", __FILE__, __LINE__);
    if (rc == BARTINT_OK) {
        const int dev0 = f->device;
        for (int k = 1; k < G; ++k) {
            bartint_forest** slot = &f->replicas[(size_t)k];
            group_post(k, [=] {
                cudaStream_t ws = group_stream(k);
                BI_CUDA(cudaStreamWaitEvent(ws, pev, 0));
                return forest_from_image(hdr->data(), hb, blob, dev0, db, ws, true, slot);
            });
        }
        rc = group_wait();
    }
    if (pev) cudaEventDestroy(pev);
    cudaFreeAsync(blob, nullptr);
    return rc;
}

// One-shot host-buffer evaluation (the R entry points predict / bartint_candidate_var / bartint_draw_means end here).
// With a device group (bartint_init) and a matrix of at least 2^16 rows per member the rows are sharded over the
// members: every member evaluates its rows on its replica of the forest, per-point outputs go straight to their place
// in the caller's arrays, and the per-draw sums are all-gathered and added in member order.
int bartint_eval(const bartint_forest* f, const double* X, int64_t N, int32_t d, uint32_t flags, const double* weights,
                 int64_t weights_len, double* draw_mean, double* point_mean, double* point_var, double* full_pred) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    BI_REQUIRE(N >= 0 && (N * d == 0 || X != nullptr), BARTINT_ERR_ARG, "X is NULL or N < 0");
    BI_REQUIRE(d == f->d, BARTINT_ERR_ARG, "point matrix has %d columns but the forest was fit on %d variables", d, f->d);
    constexpr int64_t kMinPipelined = 1 << 18;
    const int G = group_size();
    const char* gm = std::getenv("BARTINT_GROUP_MIN_ROWS");          // rows per member below which one GPU does it all
    const int64_t group_min = gm ? (int64_t)atoll(gm) : (int64_t)1 << 16;
    const bool grouped = G > 1 && group_has(f->device) && f->S > 0 && N >= group_min * G && f->image_base == nullptr;
    if (!grouped && (N < kMinPipelined || full_pred != nullptr || f->S == 0)) {      // small / full-matrix requests: single shot
        bartint_points* p = nullptr;
        int rc = bartint_points_from_host(f, X, N, d, &p);
        if (rc) return rc;
        rc = bartint_eval_points(f, p, flags, 0, f->S, weights, weights_len, draw_mean, point_mean, point_var, full_pred);
        bartint_points_destroy(p);
        return rc;
    }
    BI_REQUIRE(weights == nullptr || weights_len == 1 || weights_len == N, BARTINT_ERR_ARG,
               "weights must have length 1 or N");
    if (!grouped) return eval_shard(f, X, N, N, d, flags, weights, weights_len, nullptr, draw_mean, point_mean, point_var, nullptr);

    int rc = ensure_replicas(f);
    if (rc) return rc;
    const int S = f->S;
    const int64_t per = ((N + G - 1) / G + 1023) / 1024 * 1024;
    std::vector<double*> d_sum((size_t)G, nullptr), d_gath((size_t)G, nullptr);
    DeviceGuard guard(f->device);
    cudaStream_t st0 = nullptr;
    BI_CUDA(cudaStreamCreateWithFlags(&st0, cudaStreamNonBlocking));
    auto shard = [=, &d_sum, &d_gath](int k, cudaStream_t ws) -> int {
        const int64_t lo = std::min<int64_t>(N, (int64_t)k * per), rows = std::min<int64_t>(N, (int64_t)(k + 1) * per) - lo;
        BI_CUDA(cudaMallocAsync((void**)&d_sum[(size_t)k], sizeof(double) * (size_t)S, ws));
        BI_CUDA(cudaMallocAsync((void**)&d_gath[(size_t)k], sizeof(double) * (size_t)S * (size_t)G, ws));
        BI_CUDA(cudaStreamSynchronize(ws));
        const bool wvec = weights != nullptr && weights_len == N;
        return eval_shard(f->replicas[(size_t)k], X + lo, rows, N, d, flags, wvec ? weights + lo : weights, wvec ? rows : weights_len,
                          d_sum[(size_t)k], nullptr, point_mean ? point_mean + lo : nullptr, point_var ? point_var + lo : nullptr,
                          full_pred ? full_pred + (size_t)S * (size_t)lo : nullptr);
    };
    for (int k = 1; k < G; ++k) group_post(k, [=] { return shard(k, group_stream(k)); });
    rc = shard(0, st0);
    const int rc2 = group_wait();
    if (rc == BARTINT_OK) rc = rc2;
    double *d_tot = nullptr, *d_out = nullptr;
    if (rc == BARTINT_OK && draw_mean != nullptr) {
        std::vector<cudaStream_t> streams((size_t)G);
        streams[0] = st0;
        for (int k = 1; k < G; ++k) streams[(size_t)k] = group_stream(k);
        rc = group_allgather(d_sum.data(), d_gath.data(), (size_t)S, streams.data());
        cudaError_t e = cudaMallocAsync((void**)&d_tot, sizeof(double) * 2 * (size_t)S, st0);
        if (rc == BARTINT_OK && e != cudaSuccess) rc = cuda_fail(e, "bartint_eval group", __FILE__, __LINE__);
        d_out = d_tot ? d_tot + S : nullptr;
        if (rc == BARTINT_OK) rc = launch_reduce_draw_parts(d_gath[0], G, S, d_tot, st0);
        if (rc == BARTINT_OK)
            rc = launch_finalize_draws(d_tot, S, N, f->y_max - f->y_min, f->y_min, (flags & BARTINT_SCALED_UNITS) ? 1 : 0, d_out, st0);
        if (rc == BARTINT_OK) {
            e = cudaMemcpyAsync(draw_mean, d_out, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost, st0);
            if (e != cudaSuccess) rc = cuda_fail(e, "bartint_eval group", __FILE__, __LINE__);
        }
    }
    cudaStreamSynchronize(st0);
    for (int k = 1; k < G; ++k) {
        double *a = d_sum[(size_t)k], *b = d_gath[(size_t)k];
        group_post(k, [=] {
            cudaStreamSynchronize(group_stream(k));
            if (a) cudaFreeAsync(a, group_stream(k));
            if (b) cudaFreeAsync(b, group_stream(k));
            return (int)BARTINT_OK;
        });
    }
    group_wait();
    if (d_sum[0]) cudaFreeAsync(d_sum[0], st0);
    if (d_gath[0]) cudaFreeAsync(d_gath[0], st0);
    if (d_tot) cudaFreeAsync(d_tot, st0);
    cudaStreamSynchronize(st0);
    cudaStreamDestroy(st0);
    return rc;
}

int bartint_leaf_indices(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                         int32_t use_table_kernel, int32_t* out) {
    int rc = check_range(f, p, draw_begin, draw_end);
    if (rc) return rc;
    const int64_t total = (int64_t)(draw_end - draw_begin) * f->m * p->N;
    if (total == 0) return BARTINT_OK;
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    BI_REQUIRE(!use_table_kernel || f->table_ok, BARTINT_ERR_UNSUPPORTED,
               "the table kernel does not apply to this forest (deep trees, > 127 cuts or too many variables)");
    DeviceGuard g(f->device);
    int32_t* d_leaf = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&d_leaf, sizeof(int32_t) * (size_t)total, nullptr));
    if (use_table_kernel) {
        if (!f->gather) rc = ensure_codes(p, nullptr);
        if (rc == BARTINT_OK)
            rc = f->gather ? launch_gather_leaf(f, p, draw_begin, draw_end, d_leaf, nullptr)
                           : launch_table_leaf(f, p, draw_begin, draw_end, d_leaf, nullptr);
    } else {
        rc = eval_core(f, p, BARTINT_FORCE_WALK, draw_begin, draw_end, nullptr, nullptr, nullptr, d_leaf, nullptr, 1.0,
                       0.0, 1, nullptr);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (rc == BARTINT_OK && e != cudaSuccess) rc = cuda_fail(e, "leaf kernel", __FILE__, __LINE__);
    if (rc == BARTINT_OK) {
        e = cudaMemcpy(out, d_leaf, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(leaf)", __FILE__, __LINE__);
    }
    cudaFreeAsync(d_leaf, nullptr);
    return rc;
}

int bartint_exact_integrals_device(const bartint_forest* f, int32_t measure, const double* lo, const double* hi,
                                   int32_t n_draws_used, double* d_out, double* d_rescaled, double* d_mean_sd,
                                   void* stream) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    BI_REQUIRE(measure >= 0 && measure <= 3, BARTINT_ERR_ARG, "unknown measure %d", measure);
    BI_REQUIRE((lo == nullptr) == (hi == nullptr), BARTINT_ERR_ARG, "lo and hi must both be given or both be NULL");
    const int n_draws = (n_draws_used <= 0 || n_draws_used > f->S) ? f->S : n_draws_used;
    if (n_draws == 0) return BARTINT_OK;
    BI_REQUIRE(d_out != nullptr, BARTINT_ERR_ARG, "d_out is NULL");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    double* d_box = nullptr;
    if (lo != nullptr && f->d > 0) {
        BI_CUDA(cudaMallocAsync((void**)&d_box, sizeof(double) * 2 * (size_t)f->d, st));
        BI_CUDA(cudaMemcpyAsync(d_box, lo, sizeof(double) * (size_t)f->d, cudaMemcpyHostToDevice, st));
        BI_CUDA(cudaMemcpyAsync(d_box + f->d, hi, sizeof(double) * (size_t)f->d, cudaMemcpyHostToDevice, st));
    }
    int rc = launch_exact_integrals(f, measure, d_box, d_box ? d_box + f->d : nullptr, n_draws, d_out, st);
    if (d_box) cudaFreeAsync(d_box, st);
    if (rc == BARTINT_OK && d_mean_sd != nullptr)
        rc = launch_finalize_integrals(d_out, n_draws, f->y_min, f->y_max, d_rescaled, d_mean_sd, st);
    return rc;
}

int bartint_finalize_draws_device(const bartint_forest* f, uint32_t flags, const double* d_draw_sum, int32_t ndraws,
                                  int64_t n_points, double* d_draw_mean, double* d_mean_var, void* stream) {
    BI_REQUIRE(f != nullptr && ndraws >= 0, BARTINT_ERR_ARG, "bad arguments");
    if (ndraws == 0) return BARTINT_OK;
    BI_REQUIRE(d_draw_sum != nullptr && d_draw_mean != nullptr, BARTINT_ERR_ARG, "NULL device pointer");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int scaled = (flags & BARTINT_SCALED_UNITS) ? 1 : 0;
    int rc = launch_finalize_draws(d_draw_sum, ndraws, n_points, f->y_max - f->y_min, f->y_min, scaled, d_draw_mean, st);
    if (rc == BARTINT_OK && d_mean_var != nullptr) rc = launch_draw_summary(d_draw_mean, ndraws, d_mean_var, st);
    return rc;
}

int bartint_exact_integrals(const bartint_forest* f, int32_t measure, const double* lo, const double* hi,
                            int32_t n_draws_used, double* out) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    BI_REQUIRE(measure >= 0 && measure <= 3, BARTINT_ERR_ARG, "unknown measure %d", measure);
    BI_REQUIRE((lo == nullptr) == (hi == nullptr), BARTINT_ERR_ARG, "lo and hi must both be given or both be NULL");
    int n_draws = (n_draws_used <= 0 || n_draws_used > f->S) ? f->S : n_draws_used;
    if (n_draws == 0) return BARTINT_OK;
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    DeviceGuard g(f->device);
    const int d = f->d;
    std::vector<double> box((size_t)2 * std::max(d, 1));
    for (int j = 0; j < d; ++j) {
        box[(size_t)j] = lo ? lo[j] : 0.0;                                            // BARTInt.R:151
        box[(size_t)d + j] = hi ? hi[j] : (measure == BARTINT_MEASURE_EXPONENTIAL ? INFINITY : 1.0);   // :153-155
    }
    double *d_box = nullptr, *d_out = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&d_box, sizeof(double) * box.size(), nullptr));
    cudaError_t e = cudaMallocAsync((void**)&d_out, sizeof(double) * (size_t)n_draws, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(d_box, box.data(), sizeof(double) * box.size(), cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? launch_exact_integrals(f, measure, d_box, d_box + d, n_draws, d_out, nullptr)
                              : cuda_fail(e, "exact_integrals setup", __FILE__, __LINE__);
    if (rc == BARTINT_OK) {
        e = cudaMemcpy(out, d_out, sizeof(double) * (size_t)n_draws, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "exact_integrals", __FILE__, __LINE__);
    }
    cudaFreeAsync(d_box, nullptr);
    cudaFreeAsync(d_out, nullptr);
    return rc;
}

int bartint_finalize_integrals_device(const bartint_forest* f, const double* d_integrals_scaled, int32_t n,
                                      double* d_rescaled, double* d_mean_sd, void* stream) {
    BI_REQUIRE(f != nullptr && n >= 1, BARTINT_ERR_ARG, "bad arguments");
    BI_REQUIRE(d_integrals_scaled != nullptr && d_mean_sd != nullptr, BARTINT_ERR_ARG, "NULL device pointer");
    DeviceGuard g(f->device);
    return launch_finalize_integrals(d_integrals_scaled, n, f->y_min, f->y_max, d_rescaled, d_mean_sd, (cudaStream_t)stream);
}

int bartint_finalize_integrals(const bartint_forest* f, const double* integrals_scaled, int32_t n, double* rescaled_out,
                               double* mean_out, double* sd_out) {
    BI_REQUIRE(f != nullptr && n >= 0, BARTINT_ERR_ARG, "bad arguments");
    if (n == 0) {
        if (mean_out) *mean_out = std::numeric_limits<double>::quiet_NaN();
        if (sd_out) *sd_out = std::numeric_limits<double>::quiet_NaN();
        return BARTINT_OK;
    }
    BI_REQUIRE(integrals_scaled != nullptr, BARTINT_ERR_ARG, "integrals is NULL");
    DeviceGuard g(f->device);
    double *d_in = nullptr, *d_res = nullptr, *d_ms = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&d_in, sizeof(double) * (size_t)(2 * n + 2), nullptr));
    d_res = d_in + n;
    d_ms = d_res + n;
    cudaError_t e = cudaMemcpy(d_in, integrals_scaled, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? launch_finalize_integrals(d_in, n, f->y_min, f->y_max, d_res, d_ms, nullptr)
                              : cuda_fail(e, "finalize_integrals", __FILE__, __LINE__);
    double ms[2] = {0, 0};
    if (rc == BARTINT_OK) {
        e = cudaMemcpy(ms, d_ms, sizeof(ms), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && rescaled_out) e = cudaMemcpy(rescaled_out, d_res, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "finalize_integrals", __FILE__, __LINE__);
    }
    cudaFreeAsync(d_in, nullptr);
    if (rc == BARTINT_OK) { if (mean_out) *mean_out = ms[0]; if (sd_out) *sd_out = ms[1]; }
    return rc;
}

int bartint_last_eval_profile(const bartint_forest* f, int32_t with_moments, float* kernel_ms, int32_t* n_launches,
                              int32_t* used_table_kernel) {
    BI_REQUIRE(f != nullptr, BARTINT_ERR_ARG, "forest is NULL");
    const int slot = with_moments ? 1 : 0;
    if (n_launches) *n_launches = f->last_launches[slot];
    if (used_table_kernel) *used_table_kernel = f->last_used_table[slot];
    if (kernel_ms) {
        *kernel_ms = 0.f;
        if (f->last_timed[slot]) {
            DeviceGuard g(f->device);
            BI_CUDA(cudaEventSynchronize(f->ev1[slot]));
            BI_CUDA(cudaEventElapsedTime(kernel_ms, f->ev0[slot], f->ev1[slot]));
        }
    }
    return BARTINT_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Device image of a forest: lets one process (one GPU) export a share of the draws and replicate the finished
// sub-forest to the other ranks with a single NCCL all-gather / broadcast, instead of every rank parsing the whole
// posterior.  host header = fixed fields + draw_group_off + cut offsets/values; device blob = the device arrays,
// each padded to 256 B, in the order of kImageArrays.
// ------------------------------------------------------------------------------------------------------------
namespace {
struct ImageHeader {
    uint32_t magic, version;
    int32_t S, m, d, max_depth, max_internal, max_cuts, route, table_ok, gather, g_regs, nb, reserved;   // reserved = count_level
    int64_t n_nodes, n_leaves, n_groups, n_cuts, device_bytes;
    double y_min, y_max;
    int64_t bytes[11];       // device array sizes (0 = absent), order of image_arrays()
    int32_t draw_internal_max, pad0;     // inputs of bs_configure: the bit-sliced form is rebuilt on the device
    double max_abs_mu;
};
constexpr uint32_t kImageMagic = 0x42494E54u;   // "BINT"

inline int64_t pad256(int64_t b) { return (b + 255) / 256 * 256; }

// the device arrays of a forest with their sizes in bytes
inline void image_arrays(const bartint_forest* f, void* const** ptrs_out, int64_t sizes[11], void** slots[11], DeviceForest* d) {
    (void)ptrs_out;
    const int64_t nt = (int64_t)f->S * f->m;
    slots[0] = (void**)&d->nodes;          sizes[0] = f->n_nodes * 8;
    slots[1] = (void**)&d->tree_off;       sizes[1] = (nt + 1) * 4;
    slots[2] = (void**)&d->tree_leaf_off;  sizes[2] = (nt + 1) * 4;
    slots[3] = (void**)&d->leaf_mu;        sizes[3] = f->n_leaves * 8;
    slots[4] = (void**)&d->cut_off;        sizes[4] = ((int64_t)f->d + 1) * 4;
    slots[5] = (void**)&d->cut_val;        sizes[5] = (int64_t)f->cut_val.size() * 8;
    slots[6] = (void**)&d->node_bit;       sizes[6] = f->table_ok ? f->n_nodes : 0;
    slots[7] = (void**)&d->tree_perm;      sizes[7] = f->table_ok ? nt * 4 : 0;
    slots[8] = (void**)&d->group_tree_off; sizes[8] = f->table_ok ? (f->n_groups + 1) * 4 : 0;
    slots[9] = (void**)&d->draw_group_off; sizes[9] = f->table_ok ? ((int64_t)f->S + 1) * 4 : 0;
    slots[10] = (void**)&d->group_rec;     sizes[10] = f->table_ok ? f->n_groups * (int64_t)fast_rec_bytes(f) + 16 : 0;
}
}  // namespace

int bartint_forest_image_size(const bartint_forest* f, int64_t* host_bytes, int64_t* device_bytes) {
    BI_REQUIRE(f != nullptr && host_bytes != nullptr && device_bytes != nullptr, BARTINT_ERR_ARG, "NULL argument");
    int64_t sizes[11];
    void** slots[11];
    DeviceForest tmp = f->dev;
    image_arrays(f, nullptr, sizes, slots, &tmp);
    int64_t total = 0;
    for (int k = 0; k < 11; ++k) total += pad256(sizes[k]);
    *device_bytes = std::max<int64_t>(total, 256);
    *host_bytes = (int64_t)sizeof(ImageHeader) + ((int64_t)f->S + 1) * 4 + ((int64_t)f->d + 1) * 4 + (int64_t)f->cut_val.size() * 8;
    return BARTINT_OK;
}

int bartint_forest_pack(const bartint_forest* f, void* host_header, int64_t host_bytes, void* d_blob, int64_t device_bytes,
                        void* stream) {
    BI_REQUIRE(f != nullptr && host_header != nullptr && d_blob != nullptr, BARTINT_ERR_ARG, "NULL argument");
    int64_t hb = 0, db = 0;
    int rc = bartint_forest_image_size(f, &hb, &db);
    if (rc) return rc;
    BI_REQUIRE(host_bytes >= hb && device_bytes >= db, BARTINT_ERR_ARG, "image buffers too small (need %lld + %lld bytes)",
               (long long)hb, (long long)db);
    DeviceGuard g(f->device);
    ImageHeader h{};
    h.magic = kImageMagic; h.version = (uint32_t)bartint_version();
    h.S = f->S; h.m = f->m; h.d = f->d; h.max_depth = f->max_depth; h.max_internal = f->max_internal; h.max_cuts = f->max_cuts;
    h.route = f->route; h.table_ok = f->table_ok ? 1 : 0; h.gather = f->gather ? 1 : 0; h.g_regs = f->g_regs; h.nb = f->nb;
    h.reserved = f->count_level;          // counting level of the records (0: none)
    h.n_nodes = f->n_nodes; h.n_leaves = f->n_leaves; h.n_groups = f->n_groups; h.n_cuts = (int64_t)f->cut_val.size();
    h.device_bytes = db; h.y_min = f->y_min; h.y_max = f->y_max;
    h.draw_internal_max = f->draw_internal_max; h.max_abs_mu = f->max_abs_mu;
    void** slots[11];
    DeviceForest tmp = f->dev;
    image_arrays(f, nullptr, h.bytes, slots, &tmp);
    unsigned char* hp = (unsigned char*)host_header;
    memcpy(hp, &h, sizeof(h)); hp += sizeof(h);
    std::vector<int32_t> dgo((size_t)f->S + 1, 0);
    if (f->draw_group_off.size() == dgo.size()) dgo = f->draw_group_off;
    memcpy(hp, dgo.data(), dgo.size() * 4); hp += dgo.size() * 4;
    memcpy(hp, f->cut_off.data(), ((size_t)f->d + 1) * 4); hp += ((size_t)f->d + 1) * 4;
    if (!f->cut_val.empty()) memcpy(hp, f->cut_val.data(), f->cut_val.size() * 8);
    // the forest's uploads ran on the default stream: order the copies after them, then copy on the caller's stream
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev = nullptr;
    BI_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, nullptr);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
    int64_t off = 0;
    for (int k = 0; k < 11 && e == cudaSuccess; ++k) {
        if (h.bytes[k] > 0)
            e = cudaMemcpyAsync((unsigned char*)d_blob + off, *slots[k], (size_t)h.bytes[k], cudaMemcpyDeviceToDevice, st);
        off += pad256(h.bytes[k]);
    }
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return cuda_fail(e, "bartint_forest_pack", __FILE__, __LINE__);
    return BARTINT_OK;
}

// src_device >= 0: the blob lives on that (peer) device and is pulled over NVLink; wait: block until the image is complete
static int forest_from_image(const void* host_header, int64_t host_bytes, const void* d_blob, int src_device,
                             int64_t device_bytes, cudaStream_t st, bool wait, bartint_forest** out) {
    BI_REQUIRE(out != nullptr, BARTINT_ERR_ARG, "out is NULL");
    *out = nullptr;
    BI_REQUIRE(host_header != nullptr && d_blob != nullptr && host_bytes >= (int64_t)sizeof(ImageHeader), BARTINT_ERR_ARG,
               "bad forest image");
    ImageHeader h;
    memcpy(&h, host_header, sizeof(h));
    BI_REQUIRE(h.magic == kImageMagic && h.version == (uint32_t)bartint_version(), BARTINT_ERR_ARG,
               "not a forest image of this library version");
    BI_REQUIRE(h.S >= 0 && h.m >= 0 && h.d >= 0 && h.n_cuts >= 0 && h.device_bytes <= device_bytes, BARTINT_ERR_ARG,
               "inconsistent forest image");
    const int64_t need = (int64_t)sizeof(ImageHeader) + ((int64_t)h.S + 1) * 4 + ((int64_t)h.d + 1) * 4 + h.n_cuts * 8;
    BI_REQUIRE(host_bytes >= need, BARTINT_ERR_ARG, "forest image header truncated");
    int rc = check_device();
    if (rc) return rc;
    bartint_forest* f = new (std::nothrow) bartint_forest();
    BI_REQUIRE(f != nullptr, BARTINT_ERR_NOMEM, "out of host memory");
    f->S = h.S; f->m = h.m; f->d = h.d; f->max_depth = h.max_depth; f->max_internal = h.max_internal; f->max_cuts = h.max_cuts;
    f->route = h.route; f->table_ok = h.table_ok != 0; f->gather = h.gather != 0; f->g_regs = h.g_regs; f->nb = h.nb;
    f->count_level = h.reserved;
    f->count_mode = h.reserved != 0;
    f->n_nodes = h.n_nodes; f->n_leaves = h.n_leaves; f->n_groups = h.n_groups; f->y_min = h.y_min; f->y_max = h.y_max;
    const unsigned char* hp = (const unsigned char*)host_header + sizeof(ImageHeader);
    f->draw_group_off.resize((size_t)h.S + 1);
    memcpy(f->draw_group_off.data(), hp, ((size_t)h.S + 1) * 4); hp += ((size_t)h.S + 1) * 4;
    f->cut_off.resize((size_t)h.d + 1);
    memcpy(f->cut_off.data(), hp, ((size_t)h.d + 1) * 4); hp += ((size_t)h.d + 1) * 4;
    f->cut_val.resize((size_t)h.n_cuts);
    if (h.n_cuts) memcpy(f->cut_val.data(), hp, (size_t)h.n_cuts * 8);
    cudaError_t e = cudaGetDevice(&f->device);
    if (e == cudaSuccess) {                 // same allocator policy as finish_forest: keep freed blocks cached
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, f->device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    // one allocation holds every array (image_base); the DeviceForest pointers point into it
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&f->image_base, (size_t)std::max<int64_t>(h.device_bytes, 256), st);
    if (e == cudaSuccess) {
        // Inside a device group the image is pulled over NVLink by a KERNEL of this device: a copy-engine transfer
        // (cudaMemcpyPeerAsync) queued behind the member's own host-to-device matrix blocks and arrived when the
        // whole shard had landed -- the forest was the last thing a member received (8 GPUs: 1.4 ms of the step).
        if (src_device >= 0 && group_size() > 1 && group_peer_reads() && src_device == group_device(0)) {
            if (launch_copy_bytes(f->image_base, d_blob, h.device_bytes, st) != BARTINT_OK) e = cudaErrorUnknown;
        } else {
            e = src_device >= 0 && src_device != f->device
                    ? cudaMemcpyPeerAsync(f->image_base, f->device, d_blob, src_device, (size_t)h.device_bytes, st)
                    : cudaMemcpyAsync(f->image_base, d_blob, (size_t)h.device_bytes, cudaMemcpyDeviceToDevice, st);
        }
    }
    if (e == cudaSuccess) {
        int64_t sizes[11];
        void** slots[11];
        image_arrays(f, nullptr, sizes, slots, &f->dev);
        int64_t off = 0;
        for (int k = 0; k < 11; ++k) {
            if (sizes[k] != h.bytes[k]) { e = cudaErrorInvalidValue; break; }
            *slots[k] = h.bytes[k] > 0 ? (void*)(f->image_base + off) : nullptr;
            off += pad256(h.bytes[k]);
        }
    }
    if (e == cudaSuccess) {                 // the bit-sliced form is not in the image: rebuild it from the walk form
        f->draw_internal_max = h.draw_internal_max;
        f->max_abs_mu = h.max_abs_mu;
        bs_configure(f);
        if (launch_build_bitslice(f, st) != BARTINT_OK) e = cudaErrorUnknown;
    }
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreate(&f->ev0[k]);
        if (e == cudaSuccess) e = cudaEventCreate(&f->ev1[k]);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_last, cudaEventDisableTiming);
    // later users may run on any stream: the image must be complete before this call returns control to them
    if (e == cudaSuccess) e = cudaEventRecord(f->ev_last, st);
    if (e == cudaSuccess) { f->ev_last_set = true; if (wait) e = cudaStreamSynchronize(st); }
    if (e != cudaSuccess) {
        rc = cuda_fail(e, "bartint_forest_unpack", __FILE__, __LINE__);
        bartint_forest_destroy(f);
        return rc;
    }
    *out = f;
    return BARTINT_OK;
}

int bartint_forest_unpack(const void* host_header, int64_t host_bytes, const void* d_blob, int64_t device_bytes, void* stream,
                          bartint_forest** out) {
    return forest_from_image(host_header, host_bytes, d_blob, -1, device_bytes, (cudaStream_t)stream, true, out);
}

// ------------------------------------------------------------------------------------------------------------
// One sequential-design step, fused and pipelined over chunks of draws ("the step on either side"):
// while the GPU evaluates the draws of chunk g, the host exports chunk g+1 (string parser, leaf mu, group planner,
// H2D, table build).  Everything the device does is enqueued on one private non-blocking stream; the exporter keeps
// using the default stream, which does not synchronise with it.
// ------------------------------------------------------------------------------------------------------------
int bartint_design_step_dbarts098(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                  int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                  double y_min, double y_max, int32_t m, int32_t S, const bartint_staged* mc_arg,
                                  const bartint_staged* cand, const double* weights, int64_t weights_len,
                                  int32_t measure, int32_t n_draws_integral, int32_t n_chunks, double* integrals_scaled,
                                  double* integral_mean_sd, double* cand_var, double* draw_mean, double* draw_mean_var) {
    BI_REQUIRE(S >= 1 && m >= 1, BARTINT_ERR_ARG, "S and m must be >= 1");
    BI_REQUIRE(tree_strings != nullptr && tree_fits != nullptr, BARTINT_ERR_ARG, "NULL dbarts state");
    BI_REQUIRE(measure >= 0 && measure <= 3, BARTINT_ERR_ARG, "unknown measure %d", measure);
    BI_REQUIRE(mc_arg == nullptr || mc_arg->d == d, BARTINT_ERR_ARG, "MC point matrix has %d columns, the model %d", mc_arg ? mc_arg->d : 0, d);
    // a matrix staged on the device group: member 0 (this thread) works on shard 0 exactly like the one-GPU path, the
    // other members get every chunk's forest image and evaluate their own rows (see the chunk loop)
    const bool grouped = mc_arg != nullptr && !mc_arg->shards.empty();
    const int GM = grouped ? (int)mc_arg->shards.size() : 1;
    BI_REQUIRE(!grouped || GM == group_size(), BARTINT_ERR_ARG, "the MC matrix was staged on a device group of %d, the group now has %d", GM, group_size());
    const bartint_staged* mc = grouped ? mc_arg->shards[0] : mc_arg;
    const int64_t N_total = mc_arg ? mc_arg->N : 0;
    BI_REQUIRE(cand == nullptr || cand->d == d, BARTINT_ERR_ARG, "candidate matrix has %d columns, the model %d", cand ? cand->d : 0, d);
    BI_REQUIRE(mc == nullptr || cand == nullptr || mc->device == cand->device, BARTINT_ERR_ARG, "matrices staged on different devices");
    const int64_t C = cand ? cand->N : 0, N = mc ? mc->N : 0;
    BI_REQUIRE(weights == nullptr || weights_len == 1 || weights_len == C, BARTINT_ERR_ARG, "weights must have length 1 or C");
    BI_REQUIRE(cand_var == nullptr || cand != nullptr, BARTINT_ERR_ARG, "cand_var wanted but no candidate matrix");
    BI_REQUIRE((draw_mean == nullptr && draw_mean_var == nullptr) || mc != nullptr, BARTINT_ERR_ARG, "draw means wanted but no MC matrix");
    int rc = check_device();
    if (rc) return rc;
    const int device = mc ? mc->device : cand ? cand->device : -1;
    int cur = 0;
    BI_CUDA(cudaGetDevice(&cur));
    DeviceGuard guard(device >= 0 ? device : cur);

    // Draw chunks: with the one-pass exporter and the bit-sliced kernel a whole cfg-2 posterior is exported in about a
    // millisecond and evaluated in half of one, so pipelining chunks against each other no longer pays for its
    // per-chunk host work (tools/step_chunks.py, tools/step_sweep.py: 1 chunk 1.9-2.1 ms, 2 chunks 2.2-2.5 ms); with
    // many candidates (group records and per-point kernels, round-1 regime) two chunks still overlap export and evaluation.
    static const int env_chunks = [] { const char* e = std::getenv("BARTINT_STEP_CHUNKS"); return e ? atoi(e) : 0; }();
    const bool few_cand = (cand ? cand->N : 0) <= 2048;
    int G = n_chunks > 0 ? n_chunks : env_chunks > 0 ? env_chunks : (S >= 256 && !(few_cand && (int64_t)S * m <= 200000) ? 2 : 1);
    G = std::min(G, S);
    const int n_int = (n_draws_integral <= 0 || n_draws_integral > S) ? S : n_draws_integral;
    const bool want_int = integrals_scaled != nullptr || integral_mean_sd != nullptr;
    const bool want_cand = cand_var != nullptr && C > 0;
    const bool want_mc = (draw_mean != nullptr || draw_mean_var != nullptr) && N_total > 0;
    const bool multi = grouped && GM > 1 && want_mc;
    // Few candidates: their per-point moments come from the node-walk kernel and the per-draw sums of the MC points from
    // the bit-sliced kernel, so the group planner, the record upload and the table build are skipped altogether.
    static const bool force_plan = [] { const char* e = std::getenv("BARTINT_STEP_PLAN"); return e && e[0] == '1'; }();
    const bool light = C <= 2048 && !force_plan;

    cudaStream_t st = nullptr;
    cudaEvent_t built = nullptr;
    std::vector<bartint_forest*> forests;
    bartint_points *pm = nullptr, *pc = nullptr;
    double* d_buf = nullptr;
    int32_t* d_counts = nullptr;
    cudaStream_t bst[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t bev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    bartint_points views[8];
    int n_bst = 0;
    // the other group members: their chunk forests, point shard, per-draw sums (complete over their rows) and the
    // receive buffer of the exchange; touched only by the member's worker thread until group_wait()
    struct Member {
        std::vector<bartint_forest*> forests;
        bartint_points* pm = nullptr;
        double *d_sum = nullptr, *d_gath = nullptr;
        // a shard that arrives in row blocks is binned and evaluated block by block, like member 0's own rows: when the
        // last block lands only that block's work is left, not the whole shard's
        bartint_points views[8];
        int n_views = 0;
        double* d_block_part = nullptr;       // [n_views][S]
        cudaEvent_t done_ev = nullptr;        // the member's row of member 0's gather buffer has been written
    };
    std::vector<Member> members((size_t)GM);
    std::vector<void*> blobs;                 // forest images on member 0, one per chunk
    std::vector<cudaEvent_t> packed_ev;
    cudaEvent_t arena_ev = nullptr;           // member 0's result buffer exists (recorded on its stream)
    auto cleanup = [&]() {
        if (multi) {
            group_wait();
            for (int g = 1; g < GM; ++g) {
                Member* M = &members[(size_t)g];
                group_post(g, [M, g] {
                    cudaStreamSynchronize(group_stream(g));
                    for (bartint_forest* f : M->forests) bartint_forest_destroy(f);
                    if (M->pm) bartint_points_destroy(M->pm);
                    if (M->d_sum) cudaFreeAsync(M->d_sum, group_stream(g));
                    if (M->d_gath) cudaFreeAsync(M->d_gath, group_stream(g));
                    if (M->d_block_part) cudaFreeAsync(M->d_block_part, group_stream(g));
                    if (M->done_ev) cudaEventDestroy(M->done_ev);
                    (void)cudaGetLastError();
                    return (int)BARTINT_OK;
                });
            }
            group_wait();
        }
        for (int b = 0; b < 8; ++b)
            if (bst[b]) cudaStreamSynchronize(bst[b]);
        if (st) cudaStreamSynchronize(st);
        cudaStreamSynchronize(nullptr);
        // (streams and events belong to the per-thread StepResources cache)
        if (pm) bartint_points_destroy(pm);
        if (pc) bartint_points_destroy(pc);
        if (d_buf) cudaFreeAsync(d_buf, nullptr);
        if (d_counts) cudaFreeAsync(d_counts, nullptr);
        for (void* b : blobs) cudaFreeAsync(b, nullptr);
        for (cudaEvent_t ev : packed_ev) cudaEventDestroy(ev);
        if (arena_ev) cudaEventDestroy(arena_ev);
        for (bartint_forest* f : forests) bartint_forest_destroy(f);
        (void)cudaGetLastError();
    };
#define DS_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #call, __FILE__, __LINE__); } \
    } while (0)
#define DS_RC(call)                                    \
    do {                                               \
        rc = (call);                                   \
        if (rc != BARTINT_OK) { cleanup(); return rc; } \
    } while (0)
    // streams and events are kept per host thread and device (creating them costs ~0.4 ms per call otherwise)
    struct StepResources {
        int device = -1;
        cudaStream_t st = nullptr, bst[8] = {};
        cudaEvent_t built = nullptr, ready = nullptr, bev[8] = {};
    };
    static thread_local StepResources res;
    {
        int dev_now = 0;
        DS_CUDA(cudaGetDevice(&dev_now));
        if (res.device != dev_now) {
            res = StepResources{};          // (resources of another device are left to the driver's teardown)
            cudaError_t e = cudaStreamCreateWithFlags(&res.st, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&res.built, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&res.ready, cudaEventDisableTiming);
            for (int b = 0; b < 8 && e == cudaSuccess; ++b) {
                e = cudaStreamCreateWithFlags(&res.bst[b], cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&res.bev[b], cudaEventDisableTiming);
            }
            if (e != cudaSuccess) { res = StepResources{}; cleanup(); return cuda_fail(e, "design step streams", __FILE__, __LINE__); }
            res.device = dev_now;
        }
    }
    st = res.st;
    built = res.built;
    ready = res.ready;
    // one device buffer: integrals S | rescaled S | mean,sd 2 | draw sums S | draw means S | mean,var 2 |
    //                    candidate means G*C | candidate M2 G*C | candidate var C | weights max(len,1)
    const size_t nS = (size_t)S, nC = (size_t)std::max<int64_t>(C, 1);
    const size_t total = 4 * nS + 4 + 2 * (size_t)G * nC + nC + (size_t)std::max<int64_t>(weights_len, 1) + 8 * nS +
                         (size_t)(GM + 1) * nS;
    DS_CUDA(cudaMallocAsync((void**)&d_buf, sizeof(double) * total, st));
    if (multi) {          // the other members write into this buffer: their streams wait until it exists
        DS_CUDA(cudaEventCreateWithFlags(&arena_ev, cudaEventDisableTiming));
        DS_CUDA(cudaEventRecord(arena_ev, st));
    }
    double* d_int = d_buf;
    double* d_int_res = d_int + nS;
    double* d_int_ms = d_int_res + nS;
    double* d_sum = d_int_ms + 2;
    double* d_dmean = d_sum + nS;
    double* d_mv = d_dmean + nS;
    double* d_cmean = d_mv + 2;
    double* d_cm2 = d_cmean + (size_t)G * nC;
    double* d_cvar = d_cm2 + (size_t)G * nC;
    double* d_w = d_cvar + nC;
    double* d_block_part = d_w + (size_t)std::max<int64_t>(weights_len, 1);       // [<= 8 row blocks][S] per-draw sums
    double* d_gath = d_block_part + 8 * nS;                                       // [GM][S] the members' per-draw sums
    double* d_sum_local = d_gath + (size_t)GM * nS;                               // member 0's own rows (group runs)
    if (weights != nullptr && want_cand)
        DS_CUDA(cudaMemcpyAsync(d_w, weights, sizeof(double) * (size_t)weights_len, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> counts((size_t)G);

    static const bool timing = [] { const char* e = std::getenv("BARTINT_TIMING"); return e && e[0] == '1'; }();
    std::vector<double> t_cpu;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec; };
    const double t_begin = now_ms();
    std::vector<int32_t> edges((size_t)G + 1, 0);
    // Chunk sizes: the host export is the critical path and the evaluation of the LAST chunk is the tail nothing can
    // hide, so the chunks shrink towards the end (weights G + 0.5 - g; BARTINT_STEP_FIRST = percent of the draws in the
    // first chunk overrides it for two chunks).  (Round 1, with a 5.6 ms evaluation behind a 2 ms export, ramped UP.)
    static const int first_pct = [] { const char* e = std::getenv("BARTINT_STEP_FIRST"); return e ? atoi(e) : 0; }();
    for (int k = 1; k <= G; ++k) {       // every chunk at least one draw, the last edge exactly S
        // sum_{g<k} (G + 0.5 - g) = k (2G + 2 - k) / 2   of   G (G + 2) / 2
        int64_t ideal = (int64_t)S * ((int64_t)k * (2 * G + 2 - k)) / ((int64_t)G * (G + 2));
        if (G == 2 && k == 1 && first_pct > 0 && first_pct < 100) ideal = (int64_t)S * first_pct / 100;
        edges[(size_t)k] = (int32_t)std::min<int64_t>(std::max<int64_t>(ideal, edges[(size_t)k - 1] + 1), S - (G - k));
    }
    for (int g = 0; g < G; ++g) {
        const int32_t s0 = edges[(size_t)g], s1 = edges[(size_t)g + 1];
        counts[(size_t)g] = s1 - s0;
        bartint_forest* f = nullptr;
        if (timing) t_cpu.push_back(now_ms() - t_begin);
        DS_RC(forest_from_dbarts_impl(tree_strings + (size_t)s0 * (size_t)m, tree_fits + (size_t)s0 * (size_t)m * (size_t)n,
                                      x_train, n, d, cut_offsets, cut_values, y_min, y_max, m, s1 - s0, true, &f, light));
        forests.push_back(f);
        // the sub-forest's copies and table build are still in flight on the default stream: order `st` after them
        DS_CUDA(cudaEventRecord(built, nullptr));
        stamp(nullptr, f->device, "forest built, chunk", g);
        DS_CUDA(cudaStreamWaitEvent(st, built, 0));
        if (g == 0 && mc != nullptr) DS_CUDA(staged_submit_rest(mc));   // the first forest is on the bus: now the rest of X
        if (multi) {
            // the chunk's forest image for the other members: packed behind the uploads on the default stream, pulled
            // by every member over NVLink on its own stream, evaluated on the member's rows by its worker thread
            int64_t hb = 0, db = 0;
            DS_RC(bartint_forest_image_size(f, &hb, &db));
            auto hdr = std::make_shared<std::vector<unsigned char>>((size_t)hb);
            void* blob = nullptr;
            DS_CUDA(cudaMallocAsync(&blob, (size_t)db, nullptr));
            blobs.push_back(blob);
            DS_RC(bartint_forest_pack(f, hdr->data(), hb, blob, db, nullptr));
            cudaEvent_t pev = nullptr;
            DS_CUDA(cudaEventCreateWithFlags(&pev, cudaEventDisableTiming));
            packed_ev.push_back(pev);
            DS_CUDA(cudaEventRecord(pev, nullptr));
            const int dev0 = f->device;
            for (int k = 1; k < GM; ++k) {
                Member* M = &members[(size_t)k];
                const bartint_staged* shard = mc_arg->shards[(size_t)k];
                group_post(k, [=] {
                    cudaStream_t ws = group_stream(k);
                    BI_CUDA(cudaStreamWaitEvent(ws, pev, 0));
                    bartint_forest* fk = nullptr;
                    int r = forest_from_image(hdr->data(), hb, blob, dev0, db, ws, false, &fk);
                    if (r) return r;
                    stamp(ws, group_device(k), "forest image + bit-sliced rows ready, chunk", g);
                    M->forests.push_back(fk);
                    if (M->d_sum == nullptr) {
                        BI_CUDA(cudaMallocAsync((void**)&M->d_sum, sizeof(double) * (size_t)S, ws));
                        BI_CUDA(cudaMallocAsync((void**)&M->d_gath, sizeof(double) * (size_t)S * (size_t)GM, ws));
                        BI_CUDA(cudaMemsetAsync(M->d_sum, 0, sizeof(double) * (size_t)S, ws));
                        if (shard->N > 0 && shard->n_blocks > 1) {
                            if ((r = make_points(fk, shard->N, shard->d, ws, &M->pm))) return r;
                            BI_CUDA(cudaMallocAsync((void**)&M->d_block_part, sizeof(double) * (size_t)S * (size_t)shard->n_blocks, ws));
                            for (int b = 0; b < shard->n_blocks; ++b) {
                                const int64_t r0 = (int64_t)b * shard->block_rows, rows = std::min(shard->block_rows, shard->N - r0);
                                if (rows <= 0) break;
                                const int64_t rows_padded = (b == shard->n_blocks - 1 || r0 + shard->block_rows >= shard->N)
                                                                ? M->pm->Npad - r0 : shard->block_rows;
                                BI_CUDA(cudaStreamWaitEvent(ws, shard->block_done[b], 0));
                                if (M->pm->codes_pm) {
                                    if ((r = launch_bin_pack_points(shard->dX + r0, shard->N, rows, shard->d, fk->dev.cut_off, fk->dev.cut_val,
                                                                    fk->max_cuts, fk->route, M->pm->codes_pm + r0, rows_padded, ws))) return r;
                                    M->pm->codes_valid = false;
                                } else if ((r = launch_bin_points(shard->dX + r0, shard->N, rows, shard->d, fk->dev.cut_off, fk->dev.cut_val,
                                                                  fk->max_cuts, fk->route, M->pm->codes + r0, M->pm->Npad, ws, rows_padded))) {
                                    return r;
                                }
                                bartint_points& v = M->views[M->n_views];
                                v = *M->pm;                                   // shallow: same buffers, rows [r0, r0 + rows)
                                v.codes_valid = true;
                                v.codes = M->pm->codes + r0;
                                v.codes_pm = M->pm->codes_pm ? M->pm->codes_pm + r0 : nullptr;
                                v.N = rows;
                                ++M->n_views;
                                // (first chunk: evaluate this block before waiting for the next one)
                                if ((r = eval_core(fk, &v, 0u, 0, s1 - s0, M->d_block_part + (size_t)(M->n_views - 1) * (size_t)S + (size_t)s0,
                                                   nullptr, nullptr, nullptr, nullptr, 1.0, 0.0, 1, ws))) return r;
                                stamp(ws, group_device(k), "shard block evaluated", b);
                            }
                            return (int)BARTINT_OK;
                        }
                        if (shard->N > 0 && (r = bartint_staged_points(fk, shard, ws, &M->pm))) return r;
                    }
                    if (M->pm == nullptr) return (int)BARTINT_OK;          // an empty shard contributes zeros
                    if (M->n_views > 0) {
                        for (int b = 0; b < M->n_views; ++b)
                            if ((r = eval_core(fk, &M->views[b], 0u, 0, s1 - s0, M->d_block_part + (size_t)b * (size_t)S + (size_t)s0,
                                               nullptr, nullptr, nullptr, nullptr, 1.0, 0.0, 1, ws))) return r;
                        return (int)BARTINT_OK;
                    }
                    stamp(ws, group_device(k), "shard binned, chunk", g);
                    r = eval_core(fk, M->pm, 0u, 0, s1 - s0, M->d_sum + s0, nullptr, nullptr, nullptr, nullptr, 1.0, 0.0, 1, ws);
                    stamp(ws, group_device(k), "shard evaluated, chunk", g);
                    return r;
                });
            }
        }
        if (g == 0) {            // the cut points are the same for every chunk: bin once against the first sub-forest
            if (want_cand) DS_RC(bartint_staged_points(f, cand, st, &pc));
            if (want_mc && N > 0 && mc->n_blocks > 1) {
                // The MC matrix arrives in row blocks.  Each block gets its own stream: bin the block when it has landed,
                // then evaluate chunk after chunk of draws on it (a view of the point set) as the sub-forests become
                // available.  The GPU always has the ready (block, chunk) pairs to run, so neither the bus transfer nor
                // the export of later chunks leaves it idle.
                DS_RC(make_points(f, N, d, st, &pm));
                DS_CUDA(cudaEventRecord(ready, st));             // d_buf and the code arrays exist from here on
                for (int b = 0; b < mc->n_blocks; ++b) {
                    const int64_t r0 = (int64_t)b * mc->block_rows, rows = std::min(mc->block_rows, N - r0);
                    if (rows <= 0) break;
                    const int64_t rows_padded = (b == mc->n_blocks - 1 || r0 + mc->block_rows >= N) ? pm->Npad - r0 : mc->block_rows;
                    bst[n_bst] = res.bst[n_bst];
                    bev[n_bst] = res.bev[n_bst];
                    cudaStream_t sb = bst[n_bst];
                    DS_CUDA(cudaStreamWaitEvent(sb, ready, 0));
                    DS_CUDA(cudaStreamWaitEvent(sb, built, 0));
                    DS_CUDA(cudaStreamWaitEvent(sb, mc->block_done[b], 0));
                    if (pm->codes_pm) {      // gather forest: one pass to the packed records of this row block
                        DS_RC(launch_bin_pack_points(mc->dX + r0, N, rows, d, f->dev.cut_off, f->dev.cut_val, f->max_cuts,
                                                     f->route, pm->codes_pm + r0, rows_padded, sb));
                        pm->codes_valid = false;
                    } else {
                        DS_RC(launch_bin_points(mc->dX + r0, N, rows, d, f->dev.cut_off, f->dev.cut_val, f->max_cuts, f->route,
                                                pm->codes + r0, pm->Npad, sb, rows_padded));
                    }
                    views[n_bst] = *pm;                              // shallow: same buffers, rows [r0, r0 + rows)
                    views[n_bst].codes_valid = true;                 // (views never materialise rows; see below)
                    views[n_bst].codes = pm->codes + r0;
                    views[n_bst].codes_pm = pm->codes_pm ? pm->codes_pm + r0 : nullptr;
                    views[n_bst].N = rows;
                    ++n_bst;
                }
            } else if (want_mc && N > 0) {
                DS_RC(bartint_staged_points(f, mc, st, &pm));
            }
        }
        if (want_mc && n_bst > 0) {
            for (int b = 0; b < n_bst; ++b) {
                if (g > 0) DS_CUDA(cudaStreamWaitEvent(bst[b], built, 0));
                DS_RC(eval_core(f, &views[b], 0u, 0, s1 - s0, d_block_part + (size_t)b * nS + (size_t)s0, nullptr, nullptr,
                                nullptr, nullptr, 1.0, 0.0, 1, bst[b]));
                stamp(bst[b], f->device, "block evaluated", b);
            }
        }
        if (want_int && s0 < n_int)
            DS_RC(launch_exact_integrals(f, measure, nullptr, nullptr, std::min(s1, n_int) - s0, d_int + s0, st));
        if (want_cand)
            DS_RC(eval_core(f, pc, 0u, 0, s1 - s0, nullptr, d_cmean + (size_t)g * nC, d_cm2 + (size_t)g * nC, nullptr, nullptr,
                            1.0, 0.0, 1, st));
        if (want_mc && n_bst == 0 && pm != nullptr)
            DS_RC(eval_core(f, pm, 0u, 0, s1 - s0, d_sum + s0, nullptr, nullptr, nullptr, nullptr, 1.0, 0.0, 1, st));
        if (timing) t_cpu.push_back(now_ms() - t_begin);
    }
    const double scale = y_max - y_min;
    if (want_int) DS_RC(launch_finalize_integrals(d_int, n_int, y_min, y_max, d_int_res, d_int_ms, st));
    if (want_cand) {
        DS_CUDA(cudaMallocAsync((void**)&d_counts, sizeof(int32_t) * (size_t)G, st));
        DS_CUDA(cudaMemcpyAsync(d_counts, counts.data(), sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice, st));
        DS_RC(launch_merge_shards(d_cmean, d_cm2, G, d_counts, C, scale, y_min, 0, weights ? d_w : nullptr, weights_len,
                                  nullptr, d_cvar, st));
    }
    if (want_mc && n_bst > 0) {                                  // join the block streams; blocks added in fixed order
        for (int b = 0; b < n_bst; ++b) {
            DS_CUDA(cudaEventRecord(bev[b], bst[b]));
            DS_CUDA(cudaStreamWaitEvent(st, bev[b], 0));
        }
        DS_RC(launch_reduce_draw_parts(d_block_part, n_bst, S, d_sum, st));
    }
    if (multi) {
        // one exchange per call: every member's per-draw sums over its rows -> member 0, added there in member order.
        // Default: every member WRITES its row of member 0's gather buffer over NVLink peer memory -- the kernel that adds
        // the member's row blocks stores its result there directly -- and member 0's stream waits for one event per
        // member.  BARTINT_GROUP_EXCHANGE=nccl (or no peer access): ncclAllGather on the compute streams.
        const bool peer = group_peer_writes();
        const int dev0 = group_device(0);
        for (int k = 1; k < GM; ++k) {
            Member* M = &members[(size_t)k];
            group_post(k, [=] {           // (runs after the member's chunk jobs: one queue per member)
                cudaStream_t ws = group_stream(k);
                double* dst = peer ? d_gath + (size_t)k * nS : M->d_sum;
                if (peer) BI_CUDA(cudaStreamWaitEvent(ws, arena_ev, 0));          // the buffer exists on member 0
                if (M->n_views > 0) {
                    int r = launch_reduce_draw_parts(M->d_block_part, M->n_views, S, dst, ws);
                    if (r) return r;
                } else if (peer) {
                    BI_CUDA(cudaMemcpyPeerAsync(dst, dev0, M->d_sum, group_device(k), sizeof(double) * nS, ws));
                }
                if (peer) {
                    if (M->done_ev == nullptr) BI_CUDA(cudaEventCreateWithFlags(&M->done_ev, cudaEventDisableTiming));
                    BI_CUDA(cudaEventRecord(M->done_ev, ws));
                }
                return (int)BARTINT_OK;
            });
        }
        DS_RC(group_wait());
        if (timing) fprintf(stderr, "[bartint]   members' work enqueued by their threads at %.3f ms\n", now_ms() - t_begin);
        if (pm == nullptr) DS_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double) * nS, st));
        stamp(st, group_device(0), "member 0 ready for the exchange");
        if (peer) {
            DS_CUDA(cudaMemcpyAsync(d_gath, d_sum, sizeof(double) * nS, cudaMemcpyDeviceToDevice, st));
            for (int k = 1; k < GM; ++k) DS_CUDA(cudaStreamWaitEvent(st, members[(size_t)k].done_ev, 0));
        } else {
            DS_CUDA(cudaMemcpyAsync(d_sum_local, d_sum, sizeof(double) * nS, cudaMemcpyDeviceToDevice, st));
            std::vector<double*> send((size_t)GM), recv((size_t)GM);
            std::vector<cudaStream_t> streams((size_t)GM);
            send[0] = d_sum_local; recv[0] = d_gath; streams[0] = st;
            for (int k = 1; k < GM; ++k) { send[(size_t)k] = members[(size_t)k].d_sum; recv[(size_t)k] = members[(size_t)k].d_gath; streams[(size_t)k] = group_stream(k); }
            DS_RC(group_allgather(send.data(), recv.data(), nS, streams.data()));
        }
        stamp(st, group_device(0), "exchange done");
        DS_RC(launch_reduce_draw_parts(d_gath, GM, S, d_sum, st));
    }
    if (want_mc) {
        DS_RC(launch_finalize_draws(d_sum, S, N_total, scale, y_min, 0, d_dmean, st));
        DS_RC(launch_draw_summary(d_dmean, S, d_mv, st));
    }
    if (integrals_scaled) DS_CUDA(cudaMemcpyAsync(integrals_scaled, d_int, sizeof(double) * (size_t)n_int, cudaMemcpyDeviceToHost, st));
    if (integral_mean_sd) DS_CUDA(cudaMemcpyAsync(integral_mean_sd, d_int_ms, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    if (want_cand) DS_CUDA(cudaMemcpyAsync(cand_var, d_cvar, sizeof(double) * (size_t)C, cudaMemcpyDeviceToHost, st));
    if (want_mc && draw_mean) DS_CUDA(cudaMemcpyAsync(draw_mean, d_dmean, sizeof(double) * nS, cudaMemcpyDeviceToHost, st));
    if (want_mc && draw_mean_var) DS_CUDA(cudaMemcpyAsync(draw_mean_var, d_mv, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    if (stamps_on()) {
        timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
        fprintf(stderr, "[bartint-stamp] %.3f host: step begin\n[bartint-stamp] %.3f host: results on the host\n", std::fmod(t_begin, 1e5),
                std::fmod(1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec, 1e5));
    }
    if (timing) {
        fprintf(stderr, "[bartint] design step: %d chunks, total %.3f ms\n", G, now_ms() - t_begin);
        for (int g = 0; g < G; ++g)
            fprintf(stderr, "[bartint]   chunk %d: host export + enqueue %.3f -> %.3f ms\n", g, t_cpu[(size_t)2 * g],
                    t_cpu[(size_t)2 * g + 1]);
    }
    cleanup();
#undef DS_CUDA
#undef DS_RC
    return BARTINT_OK;
}

}  // extern "C"
