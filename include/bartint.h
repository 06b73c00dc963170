/*
 * bartint.h -- C ABI of libbartint_cuda.so, the B200 (sm_100a) implementation of the BART-Int
 * posterior-evaluation hot path.
 *
 * The reference (ImperialCollegeLondon/BART-Int) has no FFI layer: the path sits behind plain R
 * calls.  Every entry point below states the reference R expression(s) it replaces (file:line under
 * /root/reference).  INTEGRATION.md shows the .Call shim a maintainer adds on the R side.
 *
 * Conventions
 *   - plain C, all buffers caller-owned, nothing retained after return, no exceptions/longjmp
 *     across the boundary: every function returns a bartint_status (0 = OK) and
 *     bartint_last_error() gives the thread-local message;
 *   - matrices use R layout: column-major double, 32-bit ints;
 *   - "scaled units" are dbarts' internal response units  y' = (y - y_min)/(y_max - y_min) - 0.5;
 *     "original units" are what predict() returns:  (y_max - y_min) * (sum_t mu + 0.5) + y_min;
 *   - tree (draw s, tree t) has linear index  t + s*m  == R's savedTrees[t+1, s+1];
 *   - functions ending in _device take DEVICE pointers and a cudaStream_t (passed as void*) and
 *     are asynchronous on that stream; all others take HOST pointers and block;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     BARTINT_ERR_NODEVICE.
 */
#ifndef BARTINT_H
#define BARTINT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bartint_forest bartint_forest; /* flattened posterior (S draws x m trees) resident in HBM */
typedef struct bartint_points bartint_points; /* point set binned to cut-point codes, resident in HBM   */

typedef enum {
    BARTINT_OK = 0,
    BARTINT_ERR_ARG = 1,         /* NULL / bad sizes / inconsistent arrays              */
    BARTINT_ERR_PARSE = 2,       /* malformed dbarts tree string                        */
    BARTINT_ERR_CUDA = 3,        /* CUDA runtime error (message has the cudaError text)  */
    BARTINT_ERR_UNSUPPORTED = 4, /* e.g. more than 255 cut points for one variable       */
    BARTINT_ERR_NOMEM = 5,
    BARTINT_ERR_NODEVICE = 6     /* no CUDA device visible: there is no CPU path         */
} bartint_status;

/* integration measures of fillProbabilityForNode, src/BARTInt.R:46-57 */
typedef enum {
    BARTINT_MEASURE_UNIFORM = 0,            /* :46-48  box [0,1]^d                                    */
    BARTINT_MEASURE_GAUSSIAN_REFCOMPAT = 1, /* :49-52  exactly as written (quirk B3), box [0,1]^d     */
    BARTINT_MEASURE_EXPONENTIAL = 2,        /* :53-56  box [0,Inf)^d (:153-155)                        */
    BARTINT_MEASURE_GAUSSIAN = 3            /* corrected :49-52: N(0.5,1) truncated to the box, both masses with mean 0.5
                                               (NOT what the reference computes)  */
} bartint_measure;

/* flags for bartint_eval* */
#define BARTINT_SCALED_UNITS 0x1u  /* outputs stay in dbarts scaled units (no y rescale)            */
#define BARTINT_FORCE_WALK   0x2u  /* use the generic node-walk kernel even if the table kernel fits */
#define BARTINT_NO_BITSLICE  0x4u  /* per-draw sums through the per-point kernels, not the bit-sliced one (A/B timing) */

int bartint_version(void);
const char* bartint_last_error(void);
int bartint_device_count(int* n_out);
int bartint_set_device(int device);       /* forests/points are created on the current device */

/* ---------------------------------------------------------------------------------------------
 * Device group: all GPUs of the box behind ONE caller (the reference's callers are a single R process,
 * src/BARTInt.R:259-262, 312-314; src/survey_design/bartMean.R:74-82).  bartint_init(n) opens a group over devices
 * 0..n-1 (n <= 0: every visible device): peer access, one worker thread and stream per member, one NCCL communicator
 * per member (ncclCommInitAll; libnccl.so.2 is loaded with dlopen).  A matrix staged with device = BARTINT_DEVICE_GROUP
 * is split into contiguous row shards, one per member, and bartint_design_step_dbarts098 given such a matrix
 * exports the posterior once (all host threads), replicates every chunk's forest image to the other members over
 * NVLink, evaluates each member's rows there row block by row block, and gathers the members' per-draw sums on member 0
 * (peer stores into its buffer; ncclAllGather on the compute streams with BARTINT_GROUP_EXCHANGE=nccl or without peer
 * access) for the member-ordered sum; bartint_eval shards the rows of its host matrix the same way and exchanges with
 * ncclAllGather.  Not thread safe: one caller at a time (R is).
 * bartint_group_exchange(): the name of the exchange in use (for reports). */
int bartint_init(int n_gpus);
int bartint_group_size(void);             /* 1 before bartint_init */
const char* bartint_group_exchange(void);
void bartint_shutdown(void);
/* Host threads used by the exporter / group planner (OpenMP).  n <= 0: one per host core.  Returns the number in
 * effect.  Launchers that pin OMP_NUM_THREADS=1 per rank (torchrun) make the exporter serial unless this is called;
 * a multi-rank host should pass cores / ranks-per-node. */
int bartint_set_host_threads(int n);

/* ---------------------------------------------------------------------------------------------
 * Tree-state exporter.  Replaces getTree() (src/BARTInt.R:92-133) applied to every (tree, draw):
 *   treeString <- sampler$state[[1]]@savedTrees[treeNum, sampleNum]            (:110)
 *   treeFits   <- sampler$state[[1]]@savedTreeFits[, treeNum, sampleNum]       (:111)
 *   dbarts:::buildTree(strsplit(gsub("\\.", "\\. ", treeString), " "))          (:118-119)
 *   dbarts:::fillObservationsForNode / fillPlotInfoForNode  (leaf mu from fits) (:123-125)
 *   cutPoints <- dbarts:::createCutPoints(sampler)                              (:108)
 * tree_strings[t + s*m] is the dbarts 0.9-8 pre-order string ("0 49 .1 20 ..");
 * tree_fits is n x m x S column-major (scaled units); x_train is n x d column-major;
 * cut_values[cut_offsets[j] .. cut_offsets[j+1]) are variable j's ascending cut points.
 * ------------------------------------------------------------------------------------------- */
int bartint_forest_from_dbarts098(const char* const* tree_strings, const double* tree_fits,
                                  const double* x_train, int64_t n, int32_t d,
                                  const int32_t* cut_offsets, const double* cut_values,
                                  double y_min, double y_max, int32_t m, int32_t S,
                                  bartint_forest** out);

/* Pre-flattened SoA forest (what the exporter produces; also the synthetic generator's format).
 * split_var = -1 marks a leaf; left/right are node indices relative to the tree's first node. */
int bartint_forest_from_soa(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                            const int32_t* split_var, const int32_t* split_cut_idx,
                            const int32_t* left, const int32_t* right, const double* leaf_mu,
                            const int32_t* cut_offsets, const double* cut_values,
                            double y_min, double y_max, bartint_forest** out);

/* Routing rule of a split node (x = the point's value of the split variable, c = the cut point):
 *   DBARTS  left iff x <= c, NaN -> left    (dbarts; what the reference's predict() does)
 *   BART    left iff x <  c, NaN -> right   (BART::wbart / pbart, tree::bn)
 * The rule only changes how points are binned (code = #{k : x > cut_k} vs #{k : x >= cut_k}); the evaluation
 * kernels work on codes and are shared. */
#define BARTINT_ROUTE_DBARTS 0
#define BARTINT_ROUTE_BART   1
int bartint_forest_from_soa_routed(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                   const int32_t* split_var, const int32_t* split_cut_idx, const int32_t* left,
                                   const int32_t* right, const double* leaf_mu, const int32_t* cut_offsets,
                                   const double* cut_values, double y_min, double y_max, int32_t route,
                                   bartint_forest** out);
int bartint_forest_route(const bartint_forest* f);

/* Exporter for packages that expose their trees with split VALUES instead of cut indices:
 *   dbarts >= 0.9-9   fit$getTrees() / extract(fit, "trees"): one row per node in pre-order, columns sample, tree, var
 *                     (-1 = leaf), value (split value | leaf prediction); the string format of 0.9-8 that
 *                     getTree() (src/BARTInt.R:92-133) parses no longer exists there (README.md:102-106);
 *   bartMachine       extract_raw_node_data(): nested nodes with splitAttributeM, splitValue, y_pred (README.md:202).
 * value[g] is the split value of an internal node and the leaf prediction of a leaf (split_var[g] = -1).  left/right
 * are node indices relative to the tree's first node, or both NULL when the nodes of every tree are in pre-order.
 * The cut-point lists become, per variable, the sorted distinct split values of the posterior (at most 255 each);
 * bartint_forest_cutpoints reads them back.  predict = offset + leaf_scale * sum_t leaf value; route as below
 * (bartMachine and dbarts: left iff x <= split).  The R side checks (leaf_scale, offset) against the package's own
 * predict() on a few rows when the forest is built (bart-int_b200/R/bartint.R). */
int bartint_forest_from_value_splits(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset,
                                     const int32_t* split_var, const double* value, const int32_t* left,
                                     const int32_t* right, int32_t route, double leaf_scale, double offset,
                                     bartint_forest** out);
/* cut_offsets [d + 1], cut_values [cut_offsets[d]] of a forest (either may be NULL); sizes from a first call with
 * cut_values = NULL. */
int bartint_forest_cutpoints(const bartint_forest* f, int32_t* cut_offsets, double* cut_values);

/* Exporter for BART::wbart / pbart fits (the reference's README.md:202 invites swapping
 * the BART package).  trees_text = fit$treedraws$trees, a NUL-terminated string: "ndpost ntree p\n", then draw by
 * draw and tree by tree a line with the node count followed by one line "nid var cut theta" per node (heap node ids:
 * root 1, children 2 nid and 2 nid + 1; var and cut 0-based; theta = leaf value).  cut_values / cut_offsets flatten
 * fit$treedraws$cutpoints.  offset = fit$offset (wbart: mean(y.train)): predict.wbart = offset + sum_t theta, which
 * the library expresses as y_min = offset - 0.5, y_max = offset + 0.5 of the dbarts rescale.  text_len <= 0: strlen. */
int bartint_forest_from_wbart(const char* trees_text, int64_t text_len, int32_t d, const int32_t* cut_offsets,
                              const double* cut_values, double offset, bartint_forest** out);

/* Device image of a forest, for replicating it between processes (one process per GPU): each rank exports 1/G of the
 * draws with bartint_forest_from_dbarts098 on its slice, packs the finished sub-forest, and the images are exchanged
 * with one NCCL all-gather / broadcast instead of every rank parsing the whole posterior.
 * pack: host_header (bartint_forest_image_size's host_bytes) is filled at once, the device arrays are copied into
 * d_blob (device_bytes, a DEVICE pointer) asynchronously on `stream`.  unpack builds a forest on the current device
 * from such a pair (the blob is copied, not adopted; blocks until the copy is done).  An unpacked forest evaluates,
 * integrates and reports info like any other; only bartint_forest_export_soa is unsupported (no host copy). */
int bartint_forest_image_size(const bartint_forest* f, int64_t* host_bytes, int64_t* device_bytes);
int bartint_forest_pack(const bartint_forest* f, void* host_header, int64_t host_bytes, void* d_blob, int64_t device_bytes,
                        void* stream);
int bartint_forest_unpack(const void* host_header, int64_t host_bytes, const void* d_blob, int64_t device_bytes, void* stream,
                          bartint_forest** out);

/* info[0..11] = S, m, d, n_nodes, n_leaves, max_depth, max_internal_nodes_per_tree,
 *               fast_kernel_ok (0/1), n_groups, table_bits, fast kernel kind (0 none, 1 table, 2 gather),
 *               gather code words per point */
int bartint_forest_info(const bartint_forest* f, int64_t info[12]);
/* The bit-sliced sums-only kernel (kernels_bitslice.cu; used by every evaluation that asks for per-draw sums only):
 * info[0..7]  = applies (0/1), predicate rows (= total number of cut points), 128-point chunks per tile, longest draw
 *               (internal nodes), ancestor planes, fixed-point exponent Q of the leaf differences, bytes of the node
 *               rows in HBM, shared memory per CTA;
 * info[8..12] = internal nodes at depth 0, 1, 2, 3, >= 4;  info[13] = mask rows referenced by the nodes of depth >= 1
 *               (depth + 1 each: the kernel's shared-memory floor is info[13] * chunks / 8 wavefronts per tile);
 * info[14]    = the same count over the padded slice rows (what the kernel loads); info[15] = slices of 32 draws.
 * Blocks until the forest's build has finished. */
int bartint_forest_bitslice_info(const bartint_forest* f, int64_t info[16]);
/* read the canonical (pre-order) flattening back; arrays sized from info[3] */
int bartint_forest_export_soa(const bartint_forest* f, int64_t* tree_node_offset, int32_t* split_var,
                              int32_t* split_cut_idx, int32_t* left, int32_t* right, double* leaf_mu);
void bartint_forest_destroy(bartint_forest* f);

/* ---------------------------------------------------------------------------------------------
 * Point sets.  Replaces dbarts' integer cut map of x.test inside predict()
 * (call sites src/BARTInt.R:312,380; src/survey_design/bartMean.R:47,74):
 *   code[i][j] = #{k : X[i,j] > cut[j][k]}    (NaN -> 0), so  "right iff x > cut"  <=>  code > cut index.
 * X is N x d column-major with leading dimension ld >= N.
 * ------------------------------------------------------------------------------------------- */
/* Staging of a host matrix that does not depend on the fitted model (the candidate matrix / fullData of
 * bartMean.R:26 exist before bart() returns): starts the host->device copy on an internal copy stream and returns at
 * once, so the transfer overlaps the exporter's CPU work.  bartint_staged_points() then bins it for a forest (waits
 * for the copy on `stream`); the staging buffer is released by bartint_staged_free().  device = CUDA device index.
 * Matrices of >= 2^18 rows travel in four row blocks, of which only the first is put on the bus at once (copies queue
 * first-in first-out on the copy engine and a forest upload must not wait behind the whole matrix); the rest follows
 * when the first consumer asks for the data.  X therefore has to stay valid until the staged matrix has been consumed
 * (bartint_staged_points / bartint_design_step_dbarts098 returned) or freed. */
typedef struct bartint_staged bartint_staged;
#define BARTINT_DEVICE_GROUP (-1)   /* stage on the device group of bartint_init: contiguous row shards, one per member */
int bartint_stage_matrix(const double* X, int64_t N, int32_t d, int device, bartint_staged** out);
int bartint_staged_flush(bartint_staged* s);   /* put the remaining row blocks on the bus now (no forest upload is due soon) */
int bartint_staged_points(const bartint_forest* f, const bartint_staged* s, void* stream, bartint_points** out);
void bartint_staged_free(bartint_staged* s);

int bartint_points_from_host(const bartint_forest* f, const double* X, int64_t N, int32_t d,
                             bartint_points** out);
int bartint_points_from_device(const bartint_forest* f, const double* dX, int64_t N, int32_t d,
                               int64_t ld, void* stream, bartint_points** out);
/* Monte Carlo points drawn ON THE DEVICE from the integration measure and binned on the fly (the N x d fp64 matrix is
 * never materialised).  Replaces the reference's host-side sampling of the measure --
 * runif / rtnorm(mean=.5, lower=0, upper=1) / rexp at src/BARTInt.R:301,304,307 -- when the points are only needed
 * for the ensemble evaluation.  Counter-based Philox4x32-10: element (point i, variable j) depends only on
 * (seed, first_index + i, j), so shards of one point set can be generated independently on different GPUs.
 * measure: UNIFORM, GAUSSIAN / GAUSSIAN_REFCOMPAT (both: N(0.5,1) truncated to [0,1]), EXPONENTIAL.
 * dX_out: optional device buffer (N x d column-major) receiving the generated values (tests), or NULL. */
int bartint_points_generate(const bartint_forest* f, int32_t measure, int64_t N, uint64_t seed, int64_t first_index,
                            double* dX_out, void* stream, bartint_points** out);
int64_t bartint_points_count(const bartint_points* p);
int bartint_points_codes(const bartint_points* p, uint8_t* codes_out /* N x d column-major */);
void bartint_points_destroy(bartint_points* p);

/* ---------------------------------------------------------------------------------------------
 * Ensemble evaluation + fused reductions over draws [draw_begin, draw_end).  Replaces
 *   fValues <- predict(model, X)                      BARTInt.R:312,380; bartMean.R:47,74
 *   colVars(fValues) * weights                        BARTInt.R:314,386; bartMean.R:50
 *   rowMeans(pred)                                    bartMean.R:76,82
 * without materialising the S x N matrix unless full_pred is requested.  Any output may be NULL.
 *   draw_mean [draw_end-draw_begin]  rowMeans(predict(model, X))
 *   point_mean[N]                    colMeans(predict(model, X))
 *   point_var [N]                    colVars(predict(model, X)) * weights   (n-1; NaN if one draw)
 *   full_pred [(draw_end-draw_begin) x N] column-major, rows = draws (what predict() returns)
 * weights: NULL (=1), or weights_len == 1 (scalar), or weights_len == N.
 * ------------------------------------------------------------------------------------------- */
int bartint_eval_points(const bartint_forest* f, const bartint_points* p, uint32_t flags,
                        int32_t draw_begin, int32_t draw_end,
                        const double* weights, int64_t weights_len,
                        double* draw_mean, double* point_mean, double* point_var, double* full_pred);

/* one-shot convenience: bin + evaluate all draws (what the R entry points predict / bartint_candidate_var /
 * bartint_draw_means call).  With a device group (bartint_init) and at least 2^16 rows per member (BARTINT_GROUP_MIN_ROWS)
 * the rows are sharded over the GPUs: every member evaluates its rows on its own replica of the forest (pulled once over
 * NVLink and kept with the forest), per-point outputs and the slabs of full_pred go straight to their place in the
 * caller's arrays, the per-draw sums are all-gathered (ncclAllGather) and added in member order. */
int bartint_eval(const bartint_forest* f, const double* X, int64_t N, int32_t d, uint32_t flags,
                 const double* weights, int64_t weights_len,
                 double* draw_mean, double* point_mean, double* point_var, double* full_pred);

/* Device-resident variant used by the multi-GPU host layer and the benchmark: raw partials in
 * SCALED units, written to device memory, asynchronous on `stream`.
 *   d_draw_sum[draw_end-draw_begin]  sum_i sum_t mu            (divide by N, rescale = rowMeans)
 *   d_point_mean[N], d_point_m2[N]   mean and sum of squared deviations over the draw range of
 *                                    sum_t mu  (Chan-mergeable across draw shards)            */
int bartint_eval_points_device(const bartint_forest* f, const bartint_points* p, uint32_t flags,
                               int32_t draw_begin, int32_t draw_end,
                               double* d_draw_sum, double* d_point_mean, double* d_point_m2,
                               void* stream);

/* predict(model, X) kept ON THE DEVICE: d_full_pred [(draw_end - draw_begin) x N] column-major, rows = draws (R's layout
 * of the S x N matrix, BARTInt.R:312,380; bartMean.R:47,74), original units unless BARTINT_SCALED_UNITS.  Asynchronous. */
int bartint_eval_points_full_device(const bartint_forest* f, const bartint_points* p, uint32_t flags,
                                    int32_t draw_begin, int32_t draw_end, double* d_full_pred, void* stream);

/* rowMeans + the survey estimator's summary (bartMean.R:76,80,82), asynchronous, device pointers:
 *   d_draw_mean[ndraws] = rowMeans(pred) = (ymax-ymin)*(d_draw_sum/n_points + 0.5) + ymin
 *   d_mean_var[2]       = { mean(pred), sum((rowMeans - mean)^2)/(ndraws-1) }   (a VARIANCE, quirk B6)
 * d_draw_sum holds sum_i sum_t mu over n_points points (e.g. the all-reduced sums of all point shards). */
int bartint_finalize_draws_device(const bartint_forest* f, uint32_t flags, const double* d_draw_sum, int32_t ndraws,
                                  int64_t n_points, double* d_draw_mean, double* d_mean_var, void* stream);

/* out[i] = sum over r of d_rows[r * n + i], the rows added in a fixed order (rows 0, 8, 16, .. then 1, 9, .. like the tile
 * reduction of the evaluation kernels) so that every rank of a point-sharded evaluation gets bit-identical totals from
 * the all-gathered per-draw partial sums.  Asynchronous on `stream`. */
int bartint_sum_rows_device(const double* d_rows, int32_t n_rows, int32_t n, double* d_out, void* stream);

/* which(var == max(var)) on the device (src/BARTInt.R:315,387; bartMean.R:51): d_out[0] = 0-based index of the
 * first maximum, d_out[1] = number of elements equal to it; {-1, 0} if any value is NaN or n == 0.  The random tie
 * break -- and R's sample(scalar) quirk, SURVEY App. B2 -- stay with the caller. */
int bartint_argmax_ties_device(const double* d_values, int64_t n, int64_t* d_out, void* stream);

/* Batch design: the k candidates of largest value, order(values, decreasing = TRUE)[1:k] with
 * ties in index order and NaN never selected -- for k = 1 the first element of which(var == max(var)) (src/BARTInt.R:
 * 315,387; bartMean.R:51).  index_out[r] = 0-based index of the r-th largest (-1 once the non-NaN values are exhausted),
 * value_out (may be NULL) its value.  _device: device pointers, asynchronous on `stream`. */
int bartint_topk_device(const double* d_values, int64_t n, int32_t k, int64_t* d_index_out, double* d_value_out, void* stream);
int bartint_topk(const double* values, int64_t n, int32_t k, int64_t* index_out, double* value_out);

/* Chan merge of G draw-shards' per-point moments (device pointers, shard-major [g*N + i]);
 * then  var = M2/(n-1) * scale^2 * weight  and  mean = y0 + scale*(0.5 + mean)  with
 * scale = y_max - y_min, y0 = y_min (or scale = 1, y0 = -0.5 for BARTINT_SCALED_UNITS).     */
int bartint_merge_moments_device(const bartint_forest* f, uint32_t flags, int32_t G,
                                 const int32_t* shard_draws /* host, G */,
                                 const double* d_mean, const double* d_m2, int64_t N,
                                 const double* d_weights, int64_t weights_len,
                                 double* d_point_mean_out, double* d_point_var_out, void* stream);

/* Leaf ordinals (leaves numbered 0..K-1 left-to-right in pre-order) reached by
 * every point in every tree of the draw range: out[(t + (s-draw_begin)*m) * N + i].  The
 * "bit-exact against dbarts routing" check reads this.  use_table_kernel != 0 decodes the
 * ordinals from the table kernel's predicate bits instead of the node-walk kernel.          */
int bartint_leaf_indices(const bartint_forest* f, const bartint_points* p,
                         int32_t draw_begin, int32_t draw_end, int32_t use_table_kernel,
                         int32_t* out);

/* ---------------------------------------------------------------------------------------------
 * Exact leaf-probability integral.  Replaces sampleIntegrals(model, dim, measure)
 * (src/BARTInt.R:204-223 -> posteriorSum :177-201 -> singleTreeSum :135-175 ->
 * fillProbabilityForNode :32-72, terminalProbability :10-30).
 * out[s], s < n_draws_used, in SCALED units (the caller rescales, BARTInt.R:260).
 * lo/hi: NULL for the reference's box ([0,1]^d, or [0,Inf)^d for the exponential measure).
 * The reference only ever uses draws 1..ntree (quirk B1): pass n_draws_used = min(m, S) for
 * bug-compatible output, S (or <= 0) for all draws.
 * ------------------------------------------------------------------------------------------- */
int bartint_exact_integrals(const bartint_forest* f, int32_t measure, const double* lo, const double* hi,
                            int32_t n_draws_used, double* out);

/* Asynchronous device variant: d_out[n_draws] scaled integrals; if d_rescaled / d_mean_sd are non-NULL the
 * finalisation of BARTInt.R:260-262 is fused behind it (d_rescaled[n_draws], d_mean_sd[2] = {mean, sd}).
 * lo/hi are HOST pointers (d doubles each) or NULL for the reference's box. */
int bartint_exact_integrals_device(const bartint_forest* f, int32_t measure, const double* lo, const double* hi,
                                   int32_t n_draws_used, double* d_out, double* d_rescaled, double* d_mean_sd,
                                   void* stream);

/* BARTInt.R:260-262 / 290-292:  integrals <- (integrals+0.5)*(ymax-ymin)+ymin; mean; sd (n-1).
 * Runs on the device the forest lives on. rescaled_out may be NULL. */
/* device-pointer variant (asynchronous on `stream`); d_rescaled may be NULL */
int bartint_finalize_integrals_device(const bartint_forest* f, const double* d_integrals_scaled, int32_t n,
                                      double* d_rescaled, double* d_mean_sd, void* stream);
int bartint_finalize_integrals(const bartint_forest* f, const double* integrals_scaled, int32_t n,
                               double* rescaled_out, double* mean_out, double* sd_out);

/* Timing of the last evaluation on this forest: milliseconds spent in the evaluation kernel (CUDA events on
 * the launching stream), kernels launched, and which kernel ran.  Two slots are kept: with_moments = 0 for the
 * last evaluation that produced per-draw sums only, 1 for the last one with per-point moments / full / leaf
 * output.  Blocks until that kernel has finished.  *used_table_kernel: 0 node-walk kernel, 1 table / gather kernel,
 * 2 bit-sliced kernel. */
int bartint_last_eval_profile(const bartint_forest* f, int32_t with_moments, float* kernel_ms, int32_t* n_launches,
                              int32_t* used_table_kernel);

/* Number of CUDA kernels this library has launched in this process so far (all devices, all streams).
 * bench.py reports the difference across its timed region as "gpu_launches". */
int64_t bartint_kernel_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * One sequential-design step in ONE call.  Replaces the body of the reference's loop
 *   sampleIntegrals(model, dim, measure); rescale, mean, sd                     BARTInt.R:259-262 / 289-292
 *   colVars(predict(model, candidateSet)) * weights                            BARTInt.R:312-314
 *   rowMeans(predict(model, X)); mean; variance of the row means               bartMean.R:74-82
 * given the dbarts state as bartint_forest_from_dbarts098 takes it and two matrices staged with
 * bartint_stage_matrix (either may be NULL).  The draws are processed in n_chunks chunks (<= 0: automatic): while
 * the GPU evaluates one chunk, the host exports the next, so the exporter's CPU time hides behind the kernels.
 * Results are identical to the separate calls up to the summation order of the Chan merge across chunks.
 *   integrals_scaled [n_draws_integral or S]  per-draw exact integrals, dbarts scaled units (sampleIntegrals' value)
 *   integral_mean_sd [2]                      mean and sd of the rescaled integrals          (BARTInt.R:260-262)
 *   cand_var         [C]                      colVars(predict(model, candidates)) * weights
 *   draw_mean        [S]                      rowMeans(predict(model, X_mc)), original units
 *   draw_mean_var    [2]                      mean(draw_mean), var(draw_mean)                 (bartMean.R:80-82)
 * Any output may be NULL.  n_draws_integral <= 0: all draws (the reference uses the first ntree, quirk B1).
 * ------------------------------------------------------------------------------------------- */
int bartint_design_step_dbarts098(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                                  int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                                  double y_min, double y_max, int32_t m, int32_t S, const bartint_staged* mc,
                                  const bartint_staged* cand, const double* weights, int64_t weights_len,
                                  int32_t measure, int32_t n_draws_integral, int32_t n_chunks, double* integrals_scaled,
                                  double* integral_mean_sd, double* cand_var, double* draw_mean, double* draw_mean_var);

#ifdef __cplusplus
}
#endif
#endif /* BARTINT_H */
