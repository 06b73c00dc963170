/*
 * bartint_test_hooks.h -- HOST-ONLY test hooks of libbartint_cuda.so.  NOT part of the drop-in surface: nothing on the
 * R side calls them and include/bartint.h does not declare them.  They exist so that the CPU test-suite (no GPU in the
 * build container) can run the exporters and the group planner exactly as forest creation runs them and compare the
 * result with the oracle.  They are compiled into the library only with -DBARTINT_TEST_HOOKS (the in-tree build of
 * bart-int_b200/build.py sets it because tests/ needs it; a production build leaves it out and the symbols are absent).
 */
#ifndef BARTINT_TEST_HOOKS_H
#define BARTINT_TEST_HOOKS_H

#include "bartint.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Test hook, HOST ONLY (no device is touched, so it runs on a box without a GPU): canonicalises a pre-flattened
 * forest exactly as bartint_forest_from_soa does and runs the same walk-form packer and group planner that
 * forest creation runs before uploading, then copies the plan out instead of sending it to the GPU.  The CPU
 * tests interpret the records with numpy (the PRMT / IDP.4A arithmetic of the gather kernel restated bit by
 * bit) and compare with the oracle, so the planner -- bin packing, byte selectors, biases, weights, record
 * flags -- is covered without a GPU.  It is not part of the drop-in surface; nothing in R calls it.
 *   info[8] = {fast kernel applies, gather format, table bits NB, code words per point R, counting level,
 *              number of group records, descriptor words per record DW, number of leaves}
 * Caller-sized outputs: walk_nodes [2 n_nodes] (x, y words of every node record), tree_leaf_offset [S m + 1],
 * leaf_mu_out [n_leaves = (n_nodes + S m) / 2], node_bit [n_nodes], tree_perm [S m], draw_group_offset [S + 1],
 * group_tree_offset [group_capacity + 1], hdr / split [group_capacity], desc [16 group_capacity].
 * group_capacity >= S m + 2 S + 1 always suffices; a smaller one that does not is BARTINT_ERR_ARG.
 * ------------------------------------------------------------------------------------------- */
int bartint_debug_plan_soa(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                           const int32_t* split_var, const int32_t* split_cut_idx, const int32_t* left,
                           const int32_t* right, const double* leaf_mu, const int32_t* cut_offsets,
                           const double* cut_values, int64_t group_capacity, int64_t info[8], uint32_t* walk_nodes,
                           int32_t* tree_leaf_offset, double* leaf_mu_out, uint8_t* node_bit, int32_t* tree_perm,
                           int32_t* draw_group_offset, int32_t* group_tree_offset, uint32_t* hdr, uint32_t* split,
                           uint32_t* desc);

/* Test hook, HOST ONLY: the exporter of bartint_forest_from_dbarts098 (string parser, leaf values from the saved
 * fits, canonical pre-order SoA, validation) followed by the walk-form packer and the group planner, i.e. every
 * host phase of forest creation, without touching a device.  The flattened forest is copied out so the CPU tests
 * can compare the exporter with the oracle's restatement of getTree() (src/BARTInt.R:92-133); it is also what
 * tools/exporter_bench.py times.  node_capacity: size of the node arrays (any of which may be NULL);
 * *n_nodes_out receives the number of nodes; BARTINT_ERR_ARG if that exceeds node_capacity. */
int bartint_debug_export_dbarts098(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                   int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                   int32_t m, int32_t S, int64_t node_capacity, int64_t* n_nodes_out,
                                   int64_t* tree_node_offset, int32_t* split_var, int32_t* split_cut_idx,
                                   int32_t* left, int32_t* right, double* leaf_mu);

/* Test hook, HOST ONLY: the exporter of bartint_forest_from_wbart (text parser, heap ids -> pre-order SoA,
 * validation) plus the walk-form packer and the planner, without touching a device.  dims_out[3] = {S, m, n_nodes}. */
int bartint_debug_export_wbart(const char* trees_text, int64_t text_len, int32_t d, const int32_t* cut_offsets,
                               const double* cut_values, int64_t tree_capacity, int64_t node_capacity,
                               int64_t dims_out[3], int64_t* tree_node_offset, int32_t* split_var,
                               int32_t* split_cut_idx, int32_t* left, int32_t* right, double* leaf_mu);

/* Test hook, HOST ONLY: the one-pass exporter of the fused design step (parse_dbarts_walk: dbarts 0.9-8 strings + fits ->
 * walk form, no canonical SoA in between).  stats[8] = {nodes, leaves, max depth, max internal nodes per tree, most
 * internal nodes in one draw, mask-row references of the nodes at depth >= 1, bit-sliced kernel applies, root splits}.
 * walk_nodes [2 node_capacity] (x, y words), tree_node_offset / tree_leaf_offset [S m + 1], leaf_mu [leaves]. */
int bartint_debug_export_dbarts098_walk(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                        int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                        int32_t m, int32_t S, int64_t node_capacity, int64_t stats[8], uint32_t* walk_nodes,
                                        int32_t* tree_node_offset, int32_t* tree_leaf_offset, double* leaf_mu);

/* Test hook, HOST ONLY: the value-split exporter of bartint_forest_from_value_splits (cut-point lists from the distinct
 * split values, cut indices, pre-order children) plus the host phases of forest creation, without touching a device.
 * Outputs sized by the caller: node arrays [tree_node_offset[S m]], cut_offsets_out [d + 1], cut_values_out [number of
 * internal nodes] (an upper bound). */
int bartint_debug_export_value_splits(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                      const int32_t* split_var, const double* value, const int32_t* left,
                                      const int32_t* right, double leaf_scale, int64_t* tree_node_offset_out,
                                      int32_t* var_out, int32_t* cut_out, int32_t* left_out, int32_t* right_out,
                                      double* leaf_mu_out, int32_t* cut_offsets_out, double* cut_values_out);

#ifdef __cplusplus
}
#endif
#endif /* BARTINT_TEST_HOOKS_H */
