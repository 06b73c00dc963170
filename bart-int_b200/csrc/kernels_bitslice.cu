// K1b  per-draw sums by BIT-SLICED evaluation:  rowMeans(predict(model, X))  without per-point work per tree.
//
// Replaces, for evaluations that only want the per-draw sums (the Monte Carlo integral of BASELINE configs 2-4 and
// the survey estimator, src/survey_design/bartMean.R:74-82 `rowMeans(pred)`), the per-point table lookups of the
// gather / table kernels.  Identity used (pre-order trees, L(a) = value of the leftmost leaf below node a):
//
//     f_t(x) = L(root) + sum over internal nodes a of  [x reaches a  and  goes right at a] * (L(right(a)) - L(a))
//
// "goes right at a" is the predicate  code_v(x) > cut  and "reaches a" is the AND of the ancestors' predicates (negated
// for ancestors left through their left child).  For a tile of T points every predicate (variable v, cut c) is ONE
// bit vector M[v,c] of T bits, built once per CTA in shared memory from the points' uint8 cut codes.  A node then
// costs a few 128-bit loads, AND and POPC for the whole tile instead of one lookup per point, and
//
//     sum_{i in tile} f_t(x_i) = T * L(root) + sum_a delta_a * popc( AND of ancestor literals & M[v_a, c_a] ).
//
// delta_a is held as a 64-bit FIXED-POINT number (scale 2^Q chosen from max |mu|) and the products are accumulated in
// 64-bit integers, so the sum over the nodes of a draw and the points of a tile is exact and independent of the
// order in which nodes are visited; one rounding happens when the tile's total is converted to fp64.
//
// Work layout ("lanes are draws"): the nodes of 32 consecutive draws (a SLICE) are stored as ELL rows -- row k holds
// the k-th node of each of the 32 draws, 16 bytes per lane -- so a warp reads a row with one coalesced LDG.128, every
// lane accumulates the sum of ITS draw in registers, and no cross-lane reduction or atomics are needed.  Node classes
// by depth, each padded to the slice's longest draw with delta = 0 entries (draws are assigned to slices in the order
// of their row counts, so the padding is small):
//     A  root nodes            popc(M[v,c]) over the tile comes from a per-tile count table P[v,c]   (1 LDS.32)
//     B  children of the root  popc((M[parent] ^ pol) & M[own])                                   (2 x CH LDS.128)
//     C2, C3  depth 2 and 3    own mask ANDed with 2 / 3 ancestor literals, all references packed in the entry
//     D  depth >= 4            ancestor references in separate planes (rare)
// Shared-memory rows are CH x 16 bytes (CH x 128 points); lane q reads the chunks in the order c ^ (q & 7), so the
// eight lanes of an LDS.128 phase always hit eight different 16-byte bank groups whatever rows they address.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pipeline.cuh"

namespace bartint {

constexpr int BS_THREADS = 512;                 // 16 warps, one CTA per SM (the mask table fills shared memory)
constexpr size_t BS_SMEM_MAX = 227 * 1024;
constexpr int BS_SORT_MAX = 16384;              // draws are ordered by cost (rank sort, O(S^2)) up to this many

// ---------------------------------------------------------------------------------------------
// device builder: walk form -> ELL rows (run once per forest, right after the upload)
// ---------------------------------------------------------------------------------------------
// Node classes by depth.  A, B, C2, C3 entries are 16 bytes: {delta lo, delta hi, refs, refs}; D entries (depth >= 4,
// rare) keep their ancestors in separate planes of one u32 per entry.
//   A   z = predicate row
//   B   z = parent byte offset | negated << 31,  w = own byte offset
//   C2/C3  z = own | anc0 << 14 | neg0 << 28 | neg1 << 29 | neg2 << 30,  w = anc1 | anc2 << 14   (offsets in 16-byte units)
//   D   z = own byte offset, w = depth;  plane a: ancestor a's byte offset | negated << 31 (all-ones row beyond depth)
enum { BS_A = 0, BS_B = 1, BS_C2 = 2, BS_C3 = 3, BS_D = 4, BS_NCLS = 5 };

struct BsLayout {
    int64_t slice_bytes;
    int32_t off[BS_NCLS];                       // byte offsets of the class blocks inside a slice
    int32_t off_anc;                            // ancestor planes of class D
    int32_t cap_d;                              // rows of the D block (stride of an ancestor plane)
    int32_t anc_planes;
    int32_t row_bytes;                          // CH * 16
    int32_t ones_off;                           // byte offset of the all-ones mask row (row index n_cuts)
};

static inline int ru(int x, int q) { return (std::max(x, 1) + q - 1) / q * q; }

static BsLayout bs_layout_of(const bartint_forest* f) {
    BsLayout L{};
    const int m = f->m, R = f->bs_rows;
    // row capacities: a draw has at most min(2^k m, R) nodes at depth k; rounded up to the prefetch group of the class
    const int cap[BS_NCLS] = {ru(m, 8), ru(std::min(2 * m, R), 4), ru(std::min(4 * m, R), 4), ru(std::min(8 * m, R), 4),
                              f->max_depth > 4 ? ru(R, 2) : 2};
    int o = 0;
    for (int c = 0; c < BS_NCLS; ++c) { L.off[c] = o; o += 512 * cap[c]; }
    L.off_anc = o;
    L.cap_d = cap[BS_D];
    L.anc_planes = f->bs_anc;
    L.slice_bytes = (int64_t)o + 128ll * L.anc_planes * L.cap_d;
    L.row_bytes = f->bs_ch * 16;
    L.ones_off = f->bs_ncuts * L.row_bytes;
    return L;
}

struct BsCnt { unsigned short n[BS_NCLS], dmax, pad0, pad1; };     // 16 bytes
static_assert(sizeof(BsCnt) == 16, "BsCnt is stored as one 16-byte word");

__device__ __forceinline__ int bs_class_of(int depth) { return depth < 4 ? depth : BS_D; }

// one pass over a tree in pre-order; calls emit(depth, leaves seen before the node, right-child offset, predicate
// row, path) for every internal node.  path[k] = byte offset of the depth-k ancestor's row | negated << 31.
template <typename Emit>
__device__ __forceinline__ void bs_walk_tree(const uint2* __restrict__ nd, int cnt, const int32_t* __restrict__ cut_off,
                                             int row_bytes, Emit emit) {
    uint32_t path[64];
    int pend[64];           // depths of the ancestors whose right subtree is still to come
    int np = 0, depth = 0, leaves = 0;
    for (int a = 0; a < cnt; ++a) {
        const uint2 r = nd[a];
        const uint32_t var = r.x & 0xFFFFu;
        if (var != LEAF_VAR) {
            const uint32_t vc = (uint32_t)cut_off[var] + (r.x >> 16);
            emit(depth, leaves, r.y, vc, path);
            if (depth < 64) {
                path[depth] = (vc * (uint32_t)row_bytes) | 0x80000000u;     // the left subtree comes next: literal negated
                pend[np++] = depth;
            }
            ++depth;
        } else {
            ++leaves;
            if (np > 0) {
                const int dp = pend[--np];
                path[dp] &= 0x7FFFFFFFu;                                    // now in that ancestor's right subtree
                depth = dp + 1;
            }
        }
    }
}

__global__ void __launch_bounds__(128) bs_count_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                       const int32_t* __restrict__ cut_off, int64_t n_trees,
                                                       BsCnt* __restrict__ tree_cnt) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_trees) return;
    int n[BS_NCLS] = {0, 0, 0, 0, 0}, dmax = 0;
    bs_walk_tree(nodes + tree_off[k], tree_off[k + 1] - tree_off[k], cut_off, 16,
                 [&](int depth, int, uint32_t, uint32_t, const uint32_t*) {
                     const int c = bs_class_of(depth);
#pragma unroll
                     for (int q = 0; q < BS_NCLS; ++q) n[q] += (q == c);
                     if (c == BS_D) dmax = max(dmax, depth);
                 });
    BsCnt out{};
#pragma unroll
    for (int q = 0; q < BS_NCLS; ++q) out.n[q] = (unsigned short)n[q];
    out.dmax = (unsigned short)dmax;
    tree_cnt[k] = out;
}

// one thread per draw: first row of every tree inside its draw's column (per class), the draw's totals, its sort
// key (draws with similar row counts share a slice, so little of a slice is padding) and base[s] = sum_t L(root_t)
__global__ void __launch_bounds__(128) bs_draw_kernel(const BsCnt* __restrict__ tree_cnt, const int32_t* __restrict__ tree_leaf_off,
                                                      const double* __restrict__ leaf_mu, int S, int m,
                                                      BsCnt* __restrict__ tree_row, BsCnt* __restrict__ draw_tot,
                                                      unsigned long long* __restrict__ key, double* __restrict__ base) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int r[BS_NCLS] = {0, 0, 0, 0, 0}, dmax = 0;
    double b = 0.0;
    for (int t = 0; t < m; ++t) {
        const int64_t k = (int64_t)s * m + t;
        const BsCnt c = tree_cnt[k];
        BsCnt row{};
#pragma unroll
        for (int q = 0; q < BS_NCLS; ++q) { row.n[q] = (unsigned short)r[q]; r[q] += c.n[q]; }
        tree_row[k] = row;
        dmax = max(dmax, (int)c.dmax);
        b += leaf_mu[tree_leaf_off[k]];                 // leaf ordinal 0 = the leftmost leaf
    }
    BsCnt tot{};
#pragma unroll
    for (int q = 0; q < BS_NCLS; ++q) tot.n[q] = (unsigned short)r[q];
    tot.dmax = (unsigned short)dmax;
    draw_tot[s] = tot;
    base[s] = b;
    // Slices take the draws in key order and every class is padded to the slice's longest draw, so draws that share a
    // slice should have similar counts: the rare deep classes first (draws with D / C3 nodes together), then the B
    // count (pairs of values, so the C2 count still sorts inside a bucket), then C2, C3; ties by the draw index.
    const auto f12 = [](int v) { return (unsigned long long)min(v, 4095); };
    key[s] = ((unsigned long long)(r[BS_D] > 0) << 63) | ((unsigned long long)(r[BS_C3] > 0) << 62) | (f12(r[BS_B] >> 1) << 48) |
             (f12(r[BS_C2]) << 36) | (f12(r[BS_C3]) << 24) | (f12(r[BS_D]) << 16) | (unsigned long long)(s & 0xFFFF);
}

// rank of every draw by key: its slot (slice = slot / 32, lane = slot % 32).  sorted == 0: identity.
__global__ void __launch_bounds__(256) bs_rank_kernel(const unsigned long long* __restrict__ key, int S, int sorted,
                                                      int32_t* __restrict__ draw_slot, int32_t* __restrict__ slot_draw) {
    __shared__ unsigned long long s_key[256];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long mine = s < S ? key[s] : 0ull;
    int rank = s;
    if (sorted) {
        rank = 0;
        for (int j0 = 0; j0 < S; j0 += 256) {
            __syncthreads();
            s_key[threadIdx.x] = j0 + (int)threadIdx.x < S ? key[j0 + threadIdx.x] : ~0ull;
            __syncthreads();
            const int nj = min(256, S - j0);
            for (int j = 0; j < nj; ++j) {
                const unsigned long long o = s_key[j];
                rank += (o < mine) || (o == mine && j0 + j < s);
            }
        }
    }
    if (s < S) { draw_slot[s] = rank; slot_draw[rank] = s; }
}

// one warp per slice: the slice's row counts = the longest of its 32 draws
__global__ void __launch_bounds__(128) bs_hdr_kernel(const BsCnt* __restrict__ draw_tot, const int32_t* __restrict__ slot_draw,
                                                     int n_slices, int4* __restrict__ slice_hdr) {
    const int sl = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (sl >= n_slices) return;
    const int s = slot_draw[sl * 32 + lane];
    int v[BS_NCLS + 1] = {0, 0, 0, 0, 0, 0};
    if (s >= 0) {
        const BsCnt t = draw_tot[s];
#pragma unroll
        for (int q = 0; q < BS_NCLS; ++q) v[q] = t.n[q];
        v[BS_NCLS] = t.dmax;
    }
#pragma unroll
    for (int q = 0; q <= BS_NCLS; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] = max(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
    if (lane == 0) {
        slice_hdr[2 * sl] = make_int4(v[0], v[1], v[2], v[3]);
        slice_hdr[2 * sl + 1] = make_int4(v[4], v[5], 0, 0);
    }
}

__global__ void __launch_bounds__(128) bs_fill_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                      const int32_t* __restrict__ tree_leaf_off, const double* __restrict__ leaf_mu,
                                                      const int32_t* __restrict__ cut_off, int64_t n_trees, int m,
                                                      const BsCnt* __restrict__ tree_row, const int32_t* __restrict__ draw_slot,
                                                      BsLayout L, double fx_scale, unsigned char* __restrict__ pool) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_trees) return;
    const int slot = draw_slot[k / m], lane = slot & 31;
    unsigned char* __restrict__ sp = pool + (int64_t)(slot >> 5) * L.slice_bytes;
    const BsCnt r0 = tree_row[k];
    int row[BS_NCLS];
#pragma unroll
    for (int q = 0; q < BS_NCLS; ++q) row[q] = r0.n[q];
    const double* __restrict__ mu = leaf_mu + tree_leaf_off[k];
    const uint32_t rb = (uint32_t)L.row_bytes;
    bs_walk_tree(nodes + tree_off[k], tree_off[k + 1] - tree_off[k], cut_off, L.row_bytes,
                 [&](int depth, int leaves, uint32_t roff, uint32_t vc, const uint32_t* path) {
                     // leaves before the right child = leaves before the node + leaves of its left subtree (roff - 1 nodes)
                     const double delta = mu[leaves + (int)(roff >> 1)] - mu[leaves];
                     const long long fx = __double2ll_rn(delta * fx_scale);
                     const uint32_t lo = (uint32_t)(unsigned long long)fx, hi = (uint32_t)((unsigned long long)fx >> 32);
                     const int c = bs_class_of(depth);
                     uint4 e = make_uint4(lo, hi, 0u, 0u);
                     if (c == BS_A) {
                         e.z = vc;
                     } else if (c == BS_B) {
                         e.z = path[0];
                         e.w = vc * rb;
                     } else if (c == BS_C2 || c == BS_C3) {
                         const uint32_t u = rb >> 4;                        // 16-byte units per row
                         e.z = vc * u | ((path[0] & 0x7FFFFFFFu) >> 4) << 14 | (path[0] >> 31) << 28 | (path[1] >> 31) << 29;
                         e.w = (path[1] & 0x7FFFFFFFu) >> 4;
                         if (c == BS_C3) {
                             e.z |= (path[2] >> 31) << 30;
                             e.w |= ((path[2] & 0x7FFFFFFFu) >> 4) << 14;
                         }
                     } else {
                         e.z = vc * rb;
                         e.w = (uint32_t)depth;
                         uint32_t* __restrict__ anc = reinterpret_cast<uint32_t*>(sp + L.off_anc);
                         for (int a = 0; a < L.anc_planes; ++a)
                             anc[((int64_t)a * L.cap_d + row[BS_D]) * 32 + lane] = a < depth ? path[a] : (uint32_t)L.ones_off;
                     }
                     int rr = 0;
#pragma unroll
                     for (int q = 0; q < BS_NCLS; ++q)
                         if (q == c) { rr = row[q]; row[q] = rr + 1; }
                     reinterpret_cast<uint4*>(sp + L.off[c])[rr * 32 + lane] = e;
                 });
}

// Decides whether the bit-sliced kernel applies to the forest and sizes its tables (host; no CUDA call).
//   bs_ncuts  total number of cut points = number of predicate rows
//   bs_ch     128-point chunks per tile: the largest of 8 / 4 / 2 whose mask table fits shared memory
//   bs_rows   longest draw (internal nodes), bs_anc = ancestor planes of class D = deepest internal node
void bs_configure(bartint_forest* f) {
    f->bs_ok = false;
    const char* e = getenv("BARTINT_BITSLICE");
    if (e && e[0] == '0') return;
    if (f->S < 1 || f->m < 1 || f->d < 1) return;
    const int64_t ncuts = f->cut_off[(size_t)f->d];
    if (ncuts < 1 || f->max_depth < 1 || f->max_depth > 33) return;     // (no split anywhere: nothing to slice)
    if (f->draw_internal_max < 1 || f->draw_internal_max >= 65535 || f->m >= 8000) return;
    int ch = 0;
    for (int c : {8, 4, 2})
        if ((size_t)(ncuts + 1) * (size_t)(c * 16 + 4) + 64 + (size_t)f->d * (size_t)c * 128 <= BS_SMEM_MAX) { ch = c; break; }
    if (ch == 0 || (ncuts + 1) * (int64_t)ch >= 1 << 14) return;         // row offsets are 14-bit counts of 16 bytes
    f->bs_ncuts = (int32_t)ncuts;
    f->bs_ch = ch;
    f->bs_rows = f->draw_internal_max;
    f->bs_anc = std::max(4, f->max_depth - 1);
    // fixed point: |delta| <= 2 max|mu| < 2^(e+1)  ->  |delta * 2^Q| < 2^61 with Q = 60 - e
    const double dmax = 2.0 * f->max_abs_mu;
    int ex = 0;
    if (dmax > 0.0 && std::isfinite(dmax)) ex = std::ilogb(dmax);
    else if (!std::isfinite(dmax)) return;                               // NaN / Inf leaf values: keep the fp64 kernels
    f->bs_q = 60 - ex;
    if (f->bs_q > 1000 || f->bs_q < -900) return;
    f->bs_ok = true;
}

int launch_build_bitslice(bartint_forest* f, cudaStream_t st) {
    if (!f->bs_ok) return BARTINT_OK;
    const BsLayout L = bs_layout_of(f);
    const int64_t n_trees = (int64_t)f->S * f->m;
    const int n_slices = (f->S + 31) / 32;
    const size_t pool_bytes = (size_t)L.slice_bytes * (size_t)n_slices;
    BsCnt *tree_cnt = nullptr, *tree_row = nullptr, *draw_tot = nullptr;
    unsigned long long* key = nullptr;
    int32_t* draw_slot = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_pool, pool_bytes, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_slice_hdr, 2 * sizeof(int4) * (size_t)n_slices, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_base, sizeof(double) * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_slot_draw, sizeof(int32_t) * 32 * (size_t)n_slices, st));
    BI_CUDA(cudaMallocAsync((void**)&tree_cnt, sizeof(BsCnt) * (size_t)n_trees, st));
    BI_CUDA(cudaMallocAsync((void**)&tree_row, sizeof(BsCnt) * (size_t)n_trees, st));
    BI_CUDA(cudaMallocAsync((void**)&draw_tot, sizeof(BsCnt) * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&key, sizeof(unsigned long long) * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&draw_slot, sizeof(int32_t) * (size_t)f->S, st));
    BI_CUDA(cudaMemsetAsync(f->dev.bs_pool, 0, pool_bytes, st));
    BI_CUDA(cudaMemsetAsync(f->dev.bs_slot_draw, 0xFF, sizeof(int32_t) * 32 * (size_t)n_slices, st));   // -1: empty slot
    const unsigned gt = (unsigned)((n_trees + 127) / 128);
    bs_count_kernel<<<gt, 128, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.cut_off, n_trees, tree_cnt);
    bs_draw_kernel<<<(unsigned)((f->S + 127) / 128), 128, 0, st>>>(tree_cnt, f->dev.tree_leaf_off, f->dev.leaf_mu, f->S, f->m,
                                                                  tree_row, draw_tot, key, f->dev.bs_base);
    bs_rank_kernel<<<(unsigned)((f->S + 255) / 256), 256, 0, st>>>(key, f->S, f->S <= BS_SORT_MAX ? 1 : 0, draw_slot,
                                                                  f->dev.bs_slot_draw);
    bs_hdr_kernel<<<(unsigned)((n_slices * 32 + 127) / 128), 128, 0, st>>>(draw_tot, f->dev.bs_slot_draw, n_slices,
                                                                          f->dev.bs_slice_hdr);
    bs_fill_kernel<<<gt, 128, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off, f->dev.leaf_mu, f->dev.cut_off,
                                       n_trees, f->m, tree_row, draw_slot, L, std::ldexp(1.0, f->bs_q), f->dev.bs_pool);
    count_launch(5);
    for (void* q : {(void*)tree_cnt, (void*)tree_row, (void*)draw_tot, (void*)key, (void*)draw_slot}) cudaFreeAsync(q, st);
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// the evaluation kernel
// ---------------------------------------------------------------------------------------------
struct BsParams {
    const unsigned char* pool;
    const int4* slice_hdr;
    const int32_t* slot_draw;
    BsLayout L;
    const uint4* codes_pm;      // point-major packed codes (gather forests), or null
    const uint8_t* codes;       // feature-major code rows
    int64_t Npad, N;
    const int32_t* cut_off;
    int d, ncuts;
    int draw_begin, draw_end;
    int n_slices, chunk_slices;
    double inv_scale;           // 2^-Q
    double* draw_part;          // [n_tiles][draw_end - draw_begin]
};

// shared-memory layout of the evaluation kernel: [mask rows][row counts][slice counter][staged codes d x TILE]
__host__ __device__ inline size_t bs_stage_off(int n_rows, int rowb) { return (((size_t)n_rows * (size_t)(rowb + 4) + 16 + 15) & ~(size_t)15); }

__device__ __forceinline__ int popc4(const uint4& v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }

// Population count of a long run of words with fewer POPC: POPC runs on the quarter-rate XU pipe (one warp instruction
// per 8 clocks and SM sub-partition), which a row pair of class B would keep exactly as busy as the shared-memory pipe
// (32 POPC per 64 wavefronts).  A carry-save adder over two words and the running `ones` word costs two LOP3 on the ALU
// pipe and replaces two POPC by one (of the carries, weight 2): per 4 words 4 LOP3 + 2 POPC instead of 4 POPC.
struct PopAcc {
    uint32_t ones = 0u;
    int twos = 0;                           // number of set bits seen in carry words (each counts twice)
    __device__ __forceinline__ void add2(uint32_t a, uint32_t b) {
        const uint32_t carry = (a & b) | (ones & (a ^ b));      // majority(ones, a, b): one LOP3
        ones ^= a ^ b;                                          // one LOP3
        twos += __popc(carry);
    }
    __device__ __forceinline__ void add4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { add2(a, b); add2(c, d); }
    __device__ __forceinline__ int total() const { return 2 * twos + __popc(ones); }
};
__device__ __forceinline__ uint4 lds16(const unsigned char* p) { return *reinterpret_cast<const uint4*>(p); }

// Rows [0, n) of one class, G at a time: the loads of the next group are in flight while this one is processed (the
// row blocks are padded to a multiple of G with zero entries, so a group may always be loaded whole).
template <int G, typename F>
__device__ __forceinline__ void bs_rows(const uint4* __restrict__ rows, int n, F body) {
    uint4 nxt[G];
    if (n > 0) {
#pragma unroll
        for (int j = 0; j < G; ++j) nxt[j] = __ldg(rows + (size_t)j * 32);
    }
    for (int k0 = 0; k0 < n; k0 += G) {
        uint4 cur[G];
#pragma unroll
        for (int j = 0; j < G; ++j) cur[j] = nxt[j];
        if (k0 + G < n) {
#pragma unroll
            for (int j = 0; j < G; ++j) nxt[j] = __ldg(rows + (size_t)(k0 + G + j) * 32);
        }
#pragma unroll
        for (int j = 0; j < G; ++j)
            if (k0 + j < n) body(cur[j]);
    }
}

template <int CH>
__global__ void __launch_bounds__(BS_THREADS, 1) eval_bitslice_kernel(const BsParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ROWB = CH * 16;           // bytes per predicate row
    constexpr int WORDS = CH * 4;           // 32-point words per row
    constexpr int TILE = CH * 128;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_rows = prm.ncuts + 1;
    unsigned char* const s_mask = smem;                                             // [n_rows][ROWB]
    uint32_t* const s_cnt = reinterpret_cast<uint32_t*>(smem + (size_t)n_rows * ROWB);   // [n_rows] popcount of a row
    int* const s_next = reinterpret_cast<int*>(s_cnt + n_rows);                          // slice counter of phase 4
    if (tid == 0) *s_next = 0;
    const int64_t tile_base = (int64_t)blockIdx.x * TILE;

    // ---- 0-2. the mask table.  Column w (32 points) of variable j's rows belongs to ONE thread: it clears the column,
    // sets bit i of row (j, code - 1) for each of its 32 points (plain read-modify-write, no other thread touches the
    // column), then ORs the rows downwards: M[j, c] = points with code > c.  Lanes hold consecutive columns, so every
    // access is bank-conflict free, and no barrier is needed until the table is complete. ----
    // (the tile's codes are first staged feature-major in shared memory with coalesced loads: a thread then needs two
    //  LDS.128 for the 32 codes of its column instead of 32 strided byte loads from global memory)
    unsigned char* const s_codes = smem + bs_stage_off(n_rows, ROWB);                  // [d][TILE]
    if (prm.codes_pm != nullptr) {
        for (int pt = tid; pt < TILE; pt += BS_THREADS) {
            const int64_t i = tile_base + pt;
            const uint4 v = i < prm.N ? __ldg(&prm.codes_pm[i]) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < prm.d) s_codes[j * TILE + pt] = (unsigned char)(wv[j >> 2] >> (8 * (j & 3)));
        }
    } else {
        for (int q = tid; q < prm.d * (TILE / 4); q += BS_THREADS) {                   // 4 points per thread and step
            const int j = q / (TILE / 4), p4 = (q % (TILE / 4)) * 4;
            const int64_t i = tile_base + p4;                                          // rows are padded to 4096: i + 3 < Npad
            uint32_t c4 = i < prm.Npad ? __ldg(reinterpret_cast<const uint32_t*>(prm.codes + (int64_t)j * prm.Npad + i)) : 0u;
            if (i + 3 >= prm.N) {                                                      // padding rows count as code 0
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (i + b >= prm.N) c4 &= ~(0xFFu << (8 * b));
            }
            *reinterpret_cast<uint32_t*>(s_codes + j * TILE + p4) = c4;
        }
    }
    __syncthreads();
    for (int item = tid; item < prm.d * WORDS; item += BS_THREADS) {
        const int j = item / WORDS, w = item % WORDS;
        const int r0 = __ldg(&prm.cut_off[j]), r1 = __ldg(&prm.cut_off[j + 1]);
        uint32_t* const col = reinterpret_cast<uint32_t*>(s_mask) + w;
        for (int r = r0; r < r1; ++r) col[(size_t)r * WORDS] = 0u;
        const uint4 ca = *reinterpret_cast<const uint4*>(s_codes + j * TILE + w * 32);
        const uint4 cb = *reinterpret_cast<const uint4*>(s_codes + j * TILE + w * 32 + 16);
        const uint32_t cw[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t c = (cw[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            if (c != 0u) col[(size_t)(r0 + (int)c - 1) * WORDS] |= 1u << i;
        }
        uint32_t m = 0u;
        for (int r = r1 - 1; r >= r0; --r) {
            m |= col[(size_t)r * WORDS];
            col[(size_t)r * WORDS] = m;
        }
    }
    for (int i = tid; i < WORDS; i += BS_THREADS)          // the all-ones literal of unused ancestor slots
        reinterpret_cast<uint32_t*>(s_mask)[(size_t)prm.ncuts * WORDS + i] = ~0u;
    __syncthreads();
    // chunk order of this lane: the 8 lanes of an LDS.128 phase read 8 different bank groups
    uint32_t rot[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) rot[c] = (uint32_t)((c ^ (lane & (CH - 1))) * 16);
    // ---- 3. per-row popcounts over the tile (class A reads these instead of the masks) ----
    for (int r = tid; r < n_rows; r += BS_THREADS) {
        const unsigned char* row = s_mask + (size_t)r * ROWB;
        int cnt = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) cnt += popc4(lds16(row + rot[c]));
        s_cnt[r] = (uint32_t)cnt;
    }
    __syncthreads();

    // ---- 4. the slices of this CTA's chunk, one warp per slice, one lane per draw.  Slices are sorted by cost
    // (ascending), so the warps take them from a shared counter starting with the most expensive one: with a fixed
    // assignment warp 15 got the two longest of the 32 slices and the CTA waited for it with an idle data pipe
    // (0.50 -> 0.47 ms at cfg 2).  Splitting a slice into quarter units with an integer combine in shared memory was
    // tried and lost (0.51 ms): every unit pays the L2 round trip of its first row group again. ----
    const int ndraws = prm.draw_end - prm.draw_begin;
    const int sl0 = blockIdx.y * prm.chunk_slices;
    const int sl1 = min(sl0 + prm.chunk_slices, prm.n_slices);
    for (;;) {
        int take = 0;
        if (lane == 0) take = atomicAdd(s_next, 1);
        take = __shfl_sync(0xffffffffu, take, 0);
        const int sl = sl1 - 1 - take;
        if (sl < sl0) break;
        const int s = __ldg(&prm.slot_draw[sl * 32 + lane]);
        const bool wanted = s >= prm.draw_begin && s < prm.draw_end;
        if (!__any_sync(0xffffffffu, wanted)) continue;
        const int4 h0 = __ldg(&prm.slice_hdr[2 * sl]), h1 = __ldg(&prm.slice_hdr[2 * sl + 1]);
        const unsigned char* __restrict__ sp = prm.pool + (int64_t)sl * prm.L.slice_bytes;
        unsigned long long acc0 = 0ull;      // sum of (low 32 bits of delta) * count
        long long acc1 = 0ll;                // sum of (high 32 bits of delta, signed) * count
        auto add = [&](const uint4& e, int cnt) {
            acc0 += (unsigned long long)e.x * (uint32_t)cnt;
            acc1 += (long long)(int)e.y * (long long)cnt;
        };
        bs_rows<8>(reinterpret_cast<const uint4*>(sp + prm.L.off[BS_A]) + lane, h0.x,
                   [&](const uint4& e) { add(e, (int)s_cnt[e.z]); });
        bs_rows<4>(reinterpret_cast<const uint4*>(sp + prm.L.off[BS_B]) + lane, h0.y, [&](const uint4& e) {
            const uint32_t pm = (uint32_t)((int)e.z >> 31);
            const unsigned char* ap = s_mask + (e.z & 0x7FFFFFFFu);
            const unsigned char* aj = s_mask + e.w;
            PopAcc pa;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint4 mp = lds16(ap + rot[c]), mj = lds16(aj + rot[c]);
                pa.add4((mp.x ^ pm) & mj.x, (mp.y ^ pm) & mj.y, (mp.z ^ pm) & mj.z, (mp.w ^ pm) & mj.w);
            }
            add(e, pa.total());
        });
        bs_rows<4>(reinterpret_cast<const uint4*>(sp + prm.L.off[BS_C2]) + lane, h0.z, [&](const uint4& e) {
            const unsigned char* aj = s_mask + ((e.z & 0x3FFFu) << 4);
            const unsigned char* a0 = s_mask + (((e.z >> 14) & 0x3FFFu) << 4);
            const unsigned char* a1 = s_mask + ((e.w & 0x3FFFu) << 4);
            const uint32_t p0 = (uint32_t)((int)(e.z << 3) >> 31), p1 = (uint32_t)((int)(e.z << 2) >> 31);
            int cnt = 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint4 mj = lds16(aj + rot[c]), m0 = lds16(a0 + rot[c]), m1 = lds16(a1 + rot[c]);
                cnt += __popc((m0.x ^ p0) & (m1.x ^ p1) & mj.x) + __popc((m0.y ^ p0) & (m1.y ^ p1) & mj.y) +
                       __popc((m0.z ^ p0) & (m1.z ^ p1) & mj.z) + __popc((m0.w ^ p0) & (m1.w ^ p1) & mj.w);
            }
            add(e, cnt);
        });
        bs_rows<2>(reinterpret_cast<const uint4*>(sp + prm.L.off[BS_C3]) + lane, h0.w, [&](const uint4& e) {
            const unsigned char* aj = s_mask + ((e.z & 0x3FFFu) << 4);
            const unsigned char* a0 = s_mask + (((e.z >> 14) & 0x3FFFu) << 4);
            const unsigned char* a1 = s_mask + ((e.w & 0x3FFFu) << 4);
            const unsigned char* a2 = s_mask + (((e.w >> 14) & 0x3FFFu) << 4);
            const uint32_t p0 = (uint32_t)((int)(e.z << 3) >> 31), p1 = (uint32_t)((int)(e.z << 2) >> 31),
                           p2 = (uint32_t)((int)(e.z << 1) >> 31);
            int cnt = 0;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint4 mj = lds16(aj + rot[c]), m0 = lds16(a0 + rot[c]), m1 = lds16(a1 + rot[c]), m2 = lds16(a2 + rot[c]);
                cnt += __popc((m0.x ^ p0) & (m1.x ^ p1) & (m2.x ^ p2) & mj.x) + __popc((m0.y ^ p0) & (m1.y ^ p1) & (m2.y ^ p2) & mj.y) +
                       __popc((m0.z ^ p0) & (m1.z ^ p1) & (m2.z ^ p2) & mj.z) + __popc((m0.w ^ p0) & (m1.w ^ p1) & (m2.w ^ p2) & mj.w);
            }
            add(e, cnt);
        });
        if (h1.x > 0) {      // class D: depth >= 4, ancestors in planes (rare)
            const uint4* __restrict__ rows = reinterpret_cast<const uint4*>(sp + prm.L.off[BS_D]) + lane;
            const uint32_t* __restrict__ anc = reinterpret_cast<const uint32_t*>(sp + prm.L.off_anc) + lane;
            for (int k = 0; k < h1.x; ++k) {
                const uint4 e = __ldg(rows + (size_t)k * 32);
                uint4 r[CH];
                const unsigned char* aj = s_mask + e.z;
#pragma unroll
                for (int c = 0; c < CH; ++c) r[c] = lds16(aj + rot[c]);
                for (int a = 0; a < h1.y; ++a) {
                    const uint32_t w = __ldg(anc + ((size_t)a * prm.L.cap_d + k) * 32);
                    const uint32_t pm = (uint32_t)((int)w >> 31);
                    const unsigned char* ap = s_mask + (w & 0x7FFFFFFFu);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const uint4 mp = lds16(ap + rot[c]);
                        r[c].x &= mp.x ^ pm; r[c].y &= mp.y ^ pm; r[c].z &= mp.z ^ pm; r[c].w &= mp.w ^ pm;
                    }
                }
                int cnt = 0;
#pragma unroll
                for (int c = 0; c < CH; ++c) cnt += popc4(r[c]);
                add(e, cnt);
            }
        }
        if (wanted)
            prm.draw_part[(int64_t)blockIdx.x * ndraws + (s - prm.draw_begin)] =
                ((double)acc1 * 4294967296.0 + (double)acc0) * prm.inv_scale;
    }
}

size_t bs_smem_bytes(const bartint_forest* f) {
    return bs_stage_off(f->bs_ncuts + 1, f->bs_ch * 16) + (size_t)f->d * (size_t)f->bs_ch * 128;
}
int bs_tile(const bartint_forest* f) { return f->bs_ch * 128; }
int64_t bs_pool_bytes(const bartint_forest* f) { return f->bs_ok ? bs_layout_of(f).slice_bytes * ((f->S + 31) / 32) : 0; }

typedef void (*bs_kernel_t)(const BsParams);
static bs_kernel_t bs_kernel_for(int ch) {
    return ch == 8 ? eval_bitslice_kernel<8> : ch == 4 ? eval_bitslice_kernel<4> : eval_bitslice_kernel<2>;
}

int launch_eval_bitslice(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                         cudaStream_t st) {
    BI_REQUIRE(f->bs_ok && f->dev.bs_pool != nullptr, BARTINT_ERR_ARG, "internal: forest has no bit-sliced form");
    BI_REQUIRE(plan.tile == bs_tile(f), BARTINT_ERR_ARG, "internal: plan/launch tile mismatch");
    BsParams prm{};
    prm.pool = f->dev.bs_pool;
    prm.slice_hdr = f->dev.bs_slice_hdr;
    prm.slot_draw = f->dev.bs_slot_draw;
    prm.L = bs_layout_of(f);
    prm.codes_pm = p->codes_pm;
    prm.codes = p->codes;
    prm.Npad = p->Npad;
    prm.N = p->N;
    prm.cut_off = f->dev.cut_off;
    prm.d = f->d;
    prm.ncuts = f->bs_ncuts;
    prm.draw_begin = plan.draw_begin;
    prm.draw_end = plan.draw_end;
    prm.n_slices = (f->S + 31) / 32;
    prm.chunk_slices = plan.chunk_draws;          // (in slices for this kernel)
    prm.inv_scale = std::ldexp(1.0, -f->bs_q);
    prm.draw_part = draw_part;
    const size_t smem = bs_smem_bytes(f);
    bs_kernel_t k = bs_kernel_for(f->bs_ch);
    BI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.n_chunks);
    k<<<grid, BS_THREADS, smem, st>>>(prm);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// draw_sum[s] = sum over tiles of draw_part[tile][s]  +  N * base[s]   (base = sum over the draw's trees of the
// leftmost leaf value).  The tiles are added in a fixed order that depends only on their number: blocks of
// BS_RED_BLOCK tiles (grid.y), inside a block 8 interleaved partial sums like reduce_draw_parts_kernel; a second launch
// of the same kernel adds the blocks' sums when there is more than one (cfg 4: 97 656 tiles x 2000 draws = 1.56 GB of
// partials, which 63 CTAs walking all tiles read at a fraction of the HBM rate).
constexpr int BS_RED_BLOCK = 1024;
__global__ void __launch_bounds__(256) bs_reduce_kernel(const double* __restrict__ part, int n_tiles, int ndraws,
                                                        const double* __restrict__ base, double n_points,
                                                        double* __restrict__ out) {
    __shared__ double s_part[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int s = blockIdx.x * 32 + x;
    const int t0 = blockIdx.y * BS_RED_BLOCK, t1 = min(t0 + BS_RED_BLOCK, n_tiles);
    double acc = 0.0;
    if (s < ndraws)
        for (int t = t0 + y; t < t1; t += 8) acc += part[(int64_t)t * ndraws + s];
    s_part[y][x] = acc;
    __syncthreads();
    if (y == 0 && s < ndraws) {
        double tot = s_part[0][x];
#pragma unroll
        for (int k = 1; k < 8; ++k) tot += s_part[k][x];
        out[(int64_t)blockIdx.y * ndraws + s] = base != nullptr ? tot + n_points * base[s] : tot;
    }
}

int launch_bs_reduce(const bartint_forest* f, const double* draw_part, int n_tiles, int32_t draw_begin, int ndraws,
                     int64_t N, double* draw_sum, cudaStream_t st) {
    if (ndraws == 0) return BARTINT_OK;
    const double* base = f->dev.bs_base + draw_begin;
    const int n_blocks = (n_tiles + BS_RED_BLOCK - 1) / BS_RED_BLOCK;
    if (n_blocks <= 1) {
        bs_reduce_kernel<<<(ndraws + 31) / 32, 256, 0, st>>>(draw_part, n_tiles, ndraws, base, (double)N, draw_sum);
        count_launch();
    } else {
        BI_REQUIRE(n_blocks <= BS_RED_BLOCK, BARTINT_ERR_UNSUPPORTED, "more than %d tiles in one evaluation", BS_RED_BLOCK * BS_RED_BLOCK);
        double* tmp = nullptr;
        BI_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * (size_t)n_blocks * (size_t)ndraws, st));
        bs_reduce_kernel<<<dim3((unsigned)((ndraws + 31) / 32), (unsigned)n_blocks), 256, 0, st>>>(draw_part, n_tiles, ndraws, nullptr,
                                                                                                   0.0, tmp);
        bs_reduce_kernel<<<(ndraws + 31) / 32, 256, 0, st>>>(tmp, n_blocks, ndraws, base, (double)N, draw_sum);
        count_launch(2);
        cudaFreeAsync(tmp, st);
    }
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
