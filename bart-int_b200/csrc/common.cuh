// Internal definitions shared by the translation units of libbartint_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "../../include/bartint.h"

namespace bartint {

// ---------------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the C ABI; status code + thread-local message
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);          // every kernel launch of the library goes through this counter
long long launch_count();
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BI_CUDA(call)                                                            \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return ::bartint::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define BI_REQUIRE(cond, code, ...)                  \
    do {                                             \
        if (!(cond)) {                               \
            ::bartint::set_error(__VA_ARGS__);       \
            return (code);                           \
        }                                            \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device forest layout
// ---------------------------------------------------------------------------------------------
// Walk kernel node record (8 B), nodes of a tree stored in pre-order (left child = self + 1):
//   x = var | cut << 16   (var == 0xFFFF marks a leaf)
//   y = internal: offset of the right child relative to this node;  leaf: leaf ordinal in the tree
constexpr uint32_t LEAF_VAR = 0xFFFFu;

// Table kernel ("predicate-byte + joint leaf table"): consecutive trees of one draw are packed into
// groups whose internal nodes total <= NB.  One record per group, contiguous in HBM so a pipeline
// stage is one 1-D bulk (TMA) copy:
//   [ 16 B header | NB x {u32 variable, u32 bias x4} (padded to 16 B) | 2^NB doubles ]
// Predicate of node j for 4 points at once: ((codes4 + bias4) & 0x80808080) with
// bias = 127 - cut (codes <= 127), i.e. bit 7 of each byte = (code > cut) = "go right".
// Table entry [idx] = sum over the group's trees of the leaf mu selected by predicate bits idx.
constexpr int TK_THREADS = 256;        // threads per CTA
constexpr int TK_NPAD = TK_THREADS * 16;   // code rows are zero padded to a multiple of the largest CTA tile (4096 points)
constexpr int TK_TILE_MIN = TK_THREADS * 8;   // smallest CTA tile (2 quads per thread)
constexpr int TK_HDR = 16;
constexpr uint32_t TK_FLAG_END_DRAW = 1u;
// A tree with more than NB internal nodes cannot be indexed by NB predicate bits.  It gets a record of its own
// of type WALK (header only) and is walked per point from the global node array inside the same kernel (rare; the
// prior puts ~1e-4 of the trees there).  Header: {flags, first tree, number of trees, 0}.
constexpr uint32_t TK_FLAG_WALK = 2u;
// gather format only: the record holds ONE joint 2^NB table over two gather rounds; unflagged gather records are DUAL
// (two independent 16-entry tables: trees [first, first+split) -> table 1, the rest -> table 2; header.w = split).
constexpr uint32_t TK_FLAG_JOINT = 4u;
// Counting mode (gather format, opt-in BARTINT_COUNT_SPLITS=1 or 2): the trees of a draw with at most that many splits
// (and, for two splits, variables that fit one DUAL table: tg_dual_able) sit in COUNTABLE DUAL records at the end of
// the draw; the last record before them carries END_EVAL.  An evaluation that only wants per-draw sums skips the
// COUNTABLE records, closes the draw at END_EVAL, and gets those trees' sums from 1-D / 2-D cumulative code histograms
// (launch_count_single_splits); an evaluation that wants per-point results treats them like any DUAL record.
constexpr uint32_t TK_FLAG_COUNTABLE = 8u;
constexpr uint32_t TK_FLAG_END_EVAL = 16u;
__host__ __device__ constexpr int tk_tbl_off(int nb) { return TK_HDR + ((nb * 8 + 15) & ~15); }   // 16 B aligned
__host__ __device__ constexpr int tk_rec_bytes(int nb) { return tk_tbl_off(nb) + (8 << nb); }       // multiple of 16

// Gather kernel (d <= 16, NB <= 5): the thread keeps each point's codes PACKED IN REGISTERS (point-major, 4 codes per
// word, R = 2..4 words per point) and one PRMT with a per-group selector gathers the codes of up to 4 nodes of one
// point; byte-wise add of the biases, sign-replicating PRMT and one IDP.4A against the per-slot weights -(8 << bit)
// give the table byte offset directly.  Two rounds per group: round 0 gathers from words (0,1) = variables 0..7,
// round 1 from words (R-2,R-1) = variables 4(R-2) .. 4(R-2)+7.  No code loads from shared memory at all.
//   record: [ 16 B header | selA|selB<<16, biasA, biasB, wA, wB, pairsel, 0, 0 | 32 doubles (2^NB if larger) ]
//   DUAL records (default) use the table area as two 16-entry tables, one per round; JOINT records as one 2^NB table.
constexpr int TG_DESC = 32;
__host__ __device__ constexpr int tg_tbl_off() { return TK_HDR + TG_DESC; }
__host__ __device__ constexpr int tg_rec_bytes(int nb) { return tg_tbl_off() + ((8 << nb) < 256 ? 256 : (8 << nb)); }
__host__ __device__ constexpr int tg_regs_for(int d) { return d <= 8 ? 2 : (d <= 12 ? 3 : 4); }
// Can a tree whose split variables live in the code words `words` (bit r = word r) sit in one table of a DUAL
// record?  Table 1 reads words (0,1); table 2 reads (R-2,R-1), (0,R-1) or (0,1,R-1).  With R <= 3 every set of words
// qualifies; with R = 4 sets containing word 2 together with word 0 or 1 do not.  The planner (which trees may become
// COUNTABLE) and count_draw_sums_kernel (which trees it adds) must agree, hence one definition for host and device.
__host__ __device__ constexpr bool tg_dual_able(unsigned words, int R) {
    const unsigned last = 1u << (R - 1);
    return (words & ~3u) == 0u || (words & ~((1u << (R - 2)) | last)) == 0u || (words & ~(3u | last)) == 0u;
}

struct DeviceForest {
    uint2* nodes = nullptr;            // n_nodes
    int32_t* tree_off = nullptr;       // S*m+1
    int32_t* tree_leaf_off = nullptr;  // S*m+1
    double* leaf_mu = nullptr;         // n_leaves
    int32_t* cut_off = nullptr;        // d+1
    double* cut_val = nullptr;
    // table path
    uint8_t* node_bit = nullptr;       // n_nodes: predicate bit of an internal node within its group
    int32_t* tree_perm = nullptr;      // S*m: permuted position -> tree id (gather planner packs trees; else identity)
    int32_t* group_tree_off = nullptr; // n_groups+1: first (permuted) tree position of each group
    int32_t* draw_group_off = nullptr; // S+1: first group of each draw
    uint8_t* group_rec = nullptr;      // n_groups * tk_rec_bytes(nb)
    // bit-sliced form (kernels_bitslice.cu), built on the device from the walk form; never part of a forest image
    unsigned char* bs_pool = nullptr;  // ELL rows of the internal nodes, one block per slice of 32 draws
    int2* bs_slice_hdr = nullptr;      // [class][slice]: rows, deepest node (class D)
    int32_t* bs_slot_draw = nullptr;   // [class][slice * 32 + lane] -> draw, -1 = empty (per class: draws sorted by row count)
    double* bs_base = nullptr;         // S: sum over the draw's trees of the leftmost leaf value
};

}  // namespace bartint

struct bartint_forest {
    int device = 0;
    int32_t S = 0, m = 0, d = 0;
    int64_t n_nodes = 0, n_leaves = 0;
    int32_t max_depth = 0, max_internal = 0, max_cuts = 0;
    double y_min = 0, y_max = 1;
    int route = 0;             // BARTINT_ROUTE_DBARTS (left iff x <= cut) or BARTINT_ROUTE_BART (left iff x < cut)
    // canonical (pre-order) host SoA
    std::vector<int64_t> tree_off;
    std::vector<int32_t> var, cut, left, right;
    std::vector<double> mu;
    std::vector<int32_t> cut_off;
    std::vector<double> cut_val;
    // table path
    bool defer_upload_sync = false;    // forest_upload leaves its default-stream work in flight (pipelined design step)
    bool walk_only = false;            // built by parse_dbarts_walk: walk form only, no canonical host SoA (design step)
    bool no_fast_plan = false;         // skip the table / gather records (design step with few candidates: the bit-sliced
                                       // kernel gives the per-draw sums, the walk kernel the few per-point results)
    bool table_ok = false;             // a fast (table or gather) kernel applies
    bool gather = false;               // group records are in the gather format (else table format)
    bool count_mode = false;           // gather format with COUNTABLE records for the trees of <= count_level splits
    int count_level = 0;               // 1: single-split trees (1-D code histograms); 2: + two-split trees (2-D)
    int g_regs = 0;                    // gather: code words per point
    int nb = 0;
    int64_t n_groups = 0;
    // bit-sliced sums-only kernel (bs_configure): applies / predicate rows / 128-point chunks per tile / longest draw
    // (internal nodes) / ancestor planes / fixed-point exponent of the leaf differences
    bool bs_ok = false;
    int32_t bs_ncuts = 0, bs_ch = 0, bs_rows = 0, bs_anc = 0, bs_q = 0, bs_perm = 0;
    int32_t draw_internal_max = 0;     // most internal nodes in one draw
    double max_abs_mu = 0.0;           // largest |leaf value|
    int64_t bs_class_nodes[5] = {0, 0, 0, 0, 0};   // internal nodes at depth 0, 1, 2, 3, >= 4 (0 for unpacked images)
    int64_t bs_refs_total = 0;         // mask rows referenced by the nodes of depth >= 1 (depth + 1 each)
    std::vector<int32_t> draw_group_off;  // host copy, S+1
    bartint::DeviceForest dev;
    // replicas on the members of the device group (bartint_eval on a group; [0] = this forest), and the generation of
    // the group they were built for
    std::vector<bartint_forest*> replicas;
    int replica_generation = -1;
    unsigned char* image_base = nullptr;  // forests restored by bartint_forest_unpack: the one allocation behind `dev`
    // profile of the last evaluation
    // slot 0: evaluations producing per-draw sums only; slot 1: evaluations with per-point moments / full output
    cudaEvent_t ev0[2] = {nullptr, nullptr}, ev1[2] = {nullptr, nullptr};
    cudaEvent_t ev_last = nullptr;     // end of the last evaluation enqueued on a caller stream (ordering for destroy)
    mutable bool ev_last_set = false;
    mutable int last_launches[2] = {0, 0};
    mutable int last_used_table[2] = {0, 0};
    mutable bool last_timed[2] = {false, false};
};

struct bartint_points {
    int device = 0;
    int64_t N = 0, Npad = 0;
    int32_t d = 0;
    cudaStream_t stream = nullptr;   // stream the code buffer was allocated on (stream-ordered allocator)
    uint8_t* codes = nullptr;  // feature-major: codes[f * Npad + i], zero padded to Npad (multiple of TK_NPAD)
    mutable bool codes_valid = true;   // false: only codes_pm has been filled (one-pass K0); ensure_codes() fills `codes`
    uint4* codes_pm = nullptr; // point-major packed codes (16 B per point) for the gather kernel; built with the codes
};

struct bartint_staged {
    int device = 0;
    int64_t N = 0;
    int32_t d = 0;
    double* dX = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t done = nullptr;
    // large matrices travel in row blocks (multiples of 4096 rows), each with its own event, so a consumer can start
    // on the first rows while the rest is still on the bus (bartint_design_step_dbarts098)
    int n_blocks = 1;
    int64_t block_rows = 0;
    cudaEvent_t block_done[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // Only block 0 is put on the bus at once: copies queue FIFO on the copy engine, and the first forest upload of a
    // design step must not wait behind the whole matrix.  The other blocks are submitted by staged_submit_rest()
    // (first consumer, or the design step right after its first forest upload).  The host matrix stays caller-owned
    // and must remain valid until the staged matrix has been consumed or freed.
    const double* host_X = nullptr;
    int64_t host_ld = 0;             // leading dimension of host_X (rows of the caller's matrix; > N for a row shard)
    int submitted = 1;
    // A matrix staged on the DEVICE GROUP (bartint_stage_matrix with device = BARTINT_DEVICE_GROUP): contiguous row
    // shards, shard g on group member g.  The container itself holds no device memory.
    std::vector<bartint_staged*> shards;
    std::vector<int64_t> shard_lo;   // first row of every shard, then N
};


namespace bartint {

// host-side pieces (forest_host.cu)
int forest_finish_host(bartint_forest* f);      // validate canonical SoA, compute stats
struct WalkForm;
int forest_upload(bartint_forest* f, WalkForm* prebuilt = nullptr);   // build device arrays (+ table groups)
// Staging memory of the walk form: page-locked blocks from a per-thread cache, so that the host->device copies of a
// forest upload are true DMA transfers that return at once (from pageable vectors cudaMemcpyAsync stages and waits:
// 0.25-0.35 ms per draw chunk of a design step).  A block released while its copy may still be in flight is tagged with
// an event on the upload stream and not handed out again before that event has completed.  Without a CUDA device (the
// host-only test hooks) the blocks are plain heap memory.  Elements are default-initialised (no zero fill).
void* pinned_stage_alloc(size_t bytes);
void pinned_stage_free(void* p, size_t bytes);
template <typename T>
struct StageAlloc {
    using value_type = T;
    StageAlloc() = default;
    template <typename U> StageAlloc(const StageAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(pinned_stage_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) { pinned_stage_free(p, n * sizeof(T)); }
    template <typename U> void construct(U*) {}                                   // default-init: no zero fill
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(static_cast<A&&>(a)...); }
    template <typename U> bool operator==(const StageAlloc<U>&) const { return true; }
    template <typename U> bool operator!=(const StageAlloc<U>&) const { return false; }
};
template <typename T> using StageVec = std::vector<T, StageAlloc<T>>;

struct WalkForm {
    StageVec<uint2> nodes;                    // n_nodes node records (see LEAF_VAR above)
    StageVec<int32_t> tree_off, leaf_off;     // S*m+1 each: first node / first leaf of every tree
    StageVec<double> leaf_mu;                 // n_leaves, by (tree, leaf ordinal)
};
// Fast-path plan of a forest: everything forest_upload sends to the device besides the walk form.  Filled by
// plan_fast_path on the host (no CUDA call in there: the CPU tests run it through bartint_debug_plan_soa and
// interpret the records with numpy against the oracle).
struct FastPlan {
    int DW = 0;                               // descriptor words per group
    std::vector<uint8_t> node_bit;            // n_nodes: predicate bit (table format) / slot (gather format) of a node
    std::vector<int32_t> tree_perm;           // S*m: permuted position -> tree id
    std::vector<int32_t> group_tree_off;      // n_groups+1: first permuted position of each group
    std::vector<uint32_t> hdr, split, desc;   // per group: flags, trees in table 1 (DUAL), DW descriptor words
};

void pack_walk_form(const bartint_forest* f, WalkForm& W);
void plan_fast_path(bartint_forest* f, const int32_t* tree_off, FastPlan& P);
int parse_dbarts_forest(const char* const* tree_strings, const double* tree_fits, const double* x_train,
                        int64_t n, int32_t d, const int32_t* cut_offsets, const double* cut_values,
                        int32_t m, int32_t S, bartint_forest* f);
int parse_dbarts_walk(const char* const* tree_strings, const double* tree_fits, const double* x_train, int64_t n, int32_t d,
                      const int32_t* cut_offsets, const double* cut_values, int32_t m, int32_t S, bartint_forest* f,
                      WalkForm& W);
int parse_wbart_forest(const char* text, int64_t len, int32_t d, bartint_forest* f);
int value_splits_to_soa(int32_t S, int32_t m, int32_t d, const int64_t* off, const int32_t* var, const double* value,
                        const int32_t* left, const int32_t* right, double leaf_scale, bartint_forest* f);
int canonicalise_soa(int32_t S, int32_t m, int32_t d, const int64_t* tree_node_offset, const int32_t* split_var,
                     const int32_t* split_cut_idx, const int32_t* left, const int32_t* right,
                     const double* leaf_mu, bartint_forest* f);

// device group (group.cu): single-process multi-GPU.  Member 0 is driven by the calling thread; every other member
// has a worker thread (its device current, one non-blocking stream) that runs posted tasks in order.
int group_size();                                   // 1 when bartint_init has not been called
int group_generation();                             // changes with every bartint_init / bartint_shutdown
int group_device(int g);
bool group_has(int device);                         // is `device` member 0 of the group
bool group_peer_reads();                            // every member can load from member 0's memory (peer access is on)
bool group_peer_writes();                           // gather-to-root by peer stores (NVLink) instead of the NCCL exchange
cudaStream_t group_stream(int g);                   // member g's worker stream (g >= 1)
void group_post(int g, std::function<int()> fn);    // g >= 1; fn returns a bartint_status
int group_wait();                                   // all posted tasks done; first error (message copied to this thread)
// all-gather of `count` doubles per member: recv[g] receives G x count (member-major) on member g.  Enqueued on
// streams[g]; NCCL (ncclAllGather, one group call) between distinct GPUs, peer copies in the same-device test mode.
int group_allgather(double* const* send, double* const* recv, size_t count, cudaStream_t const* streams);
const char* group_exchange_name();

// kernel launchers (kernels_*.cu)
int launch_bin_points(const double* dX, int64_t ld, int64_t N, int32_t d, const int32_t* d_cut_off,
                      const double* d_cut_val, int max_cuts, int route, uint8_t* codes, int64_t Npad, cudaStream_t st,
                      int64_t rows_padded = 0);
int launch_generate_points(int measure, int64_t N, int64_t first_index, uint64_t seed, int32_t d,
                           const int32_t* d_cut_off, const double* d_cut_val, int route, uint8_t* codes, int64_t Npad,
                           double* x_out, cudaStream_t st);
struct EvalPlan;
int launch_build_tables(const bartint_forest* f, const uint32_t* d_desc, int desc_words, const uint32_t* d_hdr,
                        const uint32_t* d_split, cudaStream_t st);
int launch_bin_pack_points(const double* dX, int64_t ld, int64_t N, int32_t d, const int32_t* d_cut_off,
                           const double* d_cut_val, int max_cuts, int route, uint4* codes_pm, int64_t rows,
                           cudaStream_t st);
int launch_count_single_splits(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                               double* d_draw_sum, cudaStream_t st);
int launch_unpack_codes(const uint4* codes_pm, int64_t Npad, int d, uint8_t* codes, cudaStream_t st);
int launch_repack_codes(const uint8_t* codes, int64_t Npad, int d, uint4* codes_pm, cudaStream_t st, int64_t rows = 0);
int launch_eval_gather(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                       double* pt_mean, double* pt_m2, double* full_pred, double fp_scale, double fp_y0,
                       int fp_scaled, cudaStream_t st);
int launch_gather_leaf(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                       int32_t* leaf_out, cudaStream_t st);
void bs_configure(bartint_forest* f);                        // host: does the bit-sliced kernel apply, table sizes
int launch_build_bitslice(bartint_forest* f, cudaStream_t st);   // device: walk form -> ELL rows
int bs_tile(const bartint_forest* f);
int64_t bs_pool_bytes(const bartint_forest* f);
size_t bs_smem_bytes(const bartint_forest* f);
int64_t bs_padded_refs(const bartint_forest* f);
int launch_eval_bitslice(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                         cudaStream_t st);
int launch_bs_reduce(const bartint_forest* f, const double* draw_part, int n_tiles, int32_t draw_begin, int ndraws,
                     int64_t N, double* draw_sum, cudaStream_t st);
int gather_tile(bool mom, bool full);
int gather_resident_ctas(const bartint_forest* f, bool mom, bool full);   // occupancy query for make_plan
inline int fast_rec_bytes(const bartint_forest* f) { return f->gather ? tg_rec_bytes(f->nb) : tk_rec_bytes(f->nb); }
inline int fast_tbl_off(const bartint_forest* f) { return f->gather ? tg_tbl_off() : tk_tbl_off(f->nb); }

struct EvalPlan {
    int32_t draw_begin, draw_end;
    int n_tiles, n_chunks, chunk_draws;   // (bit-sliced kernel: chunk_draws counts slices of 32 draws)
    int tile;          // points per CTA
    bool use_bitslice; // sums-only evaluation by the bit-sliced kernel (kernels_bitslice.cu)
    bool use_table;    // a fast kernel (table or gather format) is used, else the walk kernel
    bool use_gather;   // ... and it is the gather kernel
    bool count_mode;   // sums-only evaluation of a counting-mode forest: COUNTABLE records skipped, their trees counted
};
EvalPlan make_plan(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                   int32_t draw_end, bool want_moments, bool want_full);

// draw_part: [n_tiles][ndraws]; pt_mean / pt_m2: [n_chunks][N] (may be null = no moments)
int launch_eval_walk(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                     double* pt_mean, double* pt_m2, int32_t* leaf_out, double* full_pred, double fp_scale,
                     double fp_y0, int fp_scaled, cudaStream_t st);
int launch_eval_table(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                      double* pt_mean, double* pt_m2, double* full_pred, double fp_scale, double fp_y0,
                      int fp_scaled, cudaStream_t st);
int launch_table_leaf(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                      int32_t* leaf_out, cudaStream_t st);
int launch_reduce_draw_parts(const double* draw_part, int n_tiles, int ndraws, double* draw_sum, cudaStream_t st);
// dst[0, bytes) = src[0, bytes) by an SM kernel (16-byte words; src may be peer memory): no copy engine is involved, so
// the copy cannot queue behind a host-to-device transfer in flight on the destination
int launch_copy_bytes(void* dst, const void* src, int64_t bytes, cudaStream_t st);
int launch_merge_point_parts(const double* pt_mean, const double* pt_m2, int n_chunks, int chunk_draws, int ndraws,
                             int64_t N, double* mean_out, double* m2_out, cudaStream_t st);
int launch_merge_shards(const double* mean, const double* m2, int G, const int32_t* d_counts, int64_t N,
                        double scale, double y0, int scaled, const double* w, int64_t wlen, double* mean_out,
                        double* var_out, cudaStream_t st);
int launch_finalize_draws(const double* draw_sum, int ndraws, int64_t N, double scale, double y0, int scaled,
                          double* out, cudaStream_t st);
int launch_exact_integrals(const bartint_forest* f, int measure, const double* d_lo, const double* d_hi,
                           int n_draws, double* d_out, cudaStream_t st);
int launch_argmax_ties(const double* d_v, int64_t n, int64_t* d_out, cudaStream_t st);
int launch_topk(const double* d_v, int64_t n, int k, int64_t* d_idx, double* d_val, cudaStream_t st);
int launch_draw_summary(const double* d_draw_mean, int ndraws, double* d_mean_var, cudaStream_t st);
int launch_finalize_integrals(const double* d_in, int n, double y_min, double y_max, double* d_rescaled,
                              double* d_mean_sd, cudaStream_t st);

}  // namespace bartint
