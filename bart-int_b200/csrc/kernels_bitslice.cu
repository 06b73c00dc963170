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
// Work layout ("lanes are draws"): the nodes of a draw are split into CLASSES by depth, and per class the nodes of 32
// draws (a SLICE of that class) are stored as ELL rows -- row k holds the k-th node of each of the 32 draws, 16 bytes per
// lane -- so a warp reads a row with one coalesced LDG.128 and every lane accumulates the sum of ITS draw in registers.
// A slice is padded to its longest draw with delta = 0 entries; each class assigns the draws to its slices on its own,
// in the order of that class's row counts (inside permutation blocks of 1024 draws), so the padding stays small.  A warp's
// work unit is one (class, slice); at its end every lane adds its two 64-bit integers to its draw's accumulators in
// shared memory (integer atomics: exact, so the order does not matter).
//     A  root nodes            popc(M[v,c]) over the tile comes from a per-tile count table P[v,c]   (1 LDS.32)
//     B  children of the root  popc((M[parent] ^ pol) & M[own])                                   (2 x CH LDS.128)
//     C2, C3  depth 2 and 3    own mask ANDed with 2 / 3 ancestor literals, all references packed in the entry
//     D  depth >= 4            ancestor references in separate planes (rare)
// Shared-memory rows are CH x 16 bytes (CH x 128 points); lane q reads the chunks in the order k ^ (q % CH), so the
// eight lanes of an LDS.128 phase always hit eight different 16-byte bank groups whatever rows they address.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pipeline.cuh"

namespace bartint {

constexpr int BS_THREADS = 512;                 // 16 warps, one CTA per SM (the mask table fills shared memory)
constexpr size_t BS_SMEM_MAX = 227 * 1024;

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

constexpr int BS_PERM = 1024;                   // draws are permuted inside blocks of this many (32 slices of 32) ...
constexpr int BS_PERM_SMALL = 256;              // ... or of this many when the masks leave no room for 1024 accumulators

// Every class has its OWN assignment of draws to slices (draws sorted by the class's row count inside a block of BS_PERM
// draws), so a slice pads to the longest of 32 draws with nearly equal counts.  Pool = one block per class:
// [slice][row][lane] entries of 16 bytes; class D additionally [slice][plane][row][lane] u32 ancestor references.
struct BsLayout {
    int64_t cls_off[BS_NCLS];                   // byte offset of the class block in the pool
    int32_t cap[BS_NCLS];                       // rows per slice of the class
    int64_t anc_off;                            // ancestor planes of class D
    int32_t anc_planes;
    int32_t n_slices;
    int32_t row_bytes;                          // CH * 16
    int32_t ones_off;                           // byte offset of the all-ones mask row (row index n_cuts)
    int64_t total_bytes;
};

static inline int ru(int x, int q) { return (std::max(x, 1) + q - 1) / q * q; }

static BsLayout bs_layout_of(const bartint_forest* f) {
    BsLayout L{};
    const int m = f->m, R = f->bs_rows;
    // row capacities: a draw has at most min(2^k m, R) nodes at depth k; rounded up to the prefetch group of the class
    const int cap[BS_NCLS] = {ru(m, 8), ru(std::min(2 * m, R), 4), ru(std::min(4 * m, R), 4), ru(std::min(8 * m, R), 4),
                              f->max_depth > 4 ? ru(R, 2) : 2};
    L.n_slices = (f->S + 31) / 32;
    int64_t o = 0;
    for (int c = 0; c < BS_NCLS; ++c) { L.cls_off[c] = o; L.cap[c] = cap[c]; o += 512ll * cap[c] * L.n_slices; }
    L.anc_off = o;
    L.anc_planes = f->bs_anc;
    o += 128ll * L.anc_planes * cap[BS_D] * L.n_slices;
    L.total_bytes = o;
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

// one thread per draw: first row of every tree inside its draw's column (per class), the draw's totals, one sort key per
// class (row count, then the draw index) and base[s] = sum_t L(root_t)
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
#pragma unroll
    for (int q = 0; q < BS_NCLS; ++q)
        key[(size_t)q * S + s] = ((unsigned long long)r[q] << 32) | ((unsigned long long)(q == BS_D ? dmax : 0) << 20) |
                                 (unsigned long long)(s & 0xFFFFF);
}

// Slot of every draw in every class (blockIdx.y = class): its rank by key among the draws of its block of BS_PERM draws;
// slice = slot / 32, lane = slot % 32.
__global__ void __launch_bounds__(256) bs_rank_kernel(const unsigned long long* __restrict__ key, int S,
                                                      int32_t* __restrict__ draw_slot, int32_t* __restrict__ slot_draw, int n_slots, int perm) {
    __shared__ unsigned long long s_key[256];
    const int c = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long* kc = key + (size_t)c * S;
    const unsigned long long mine = s < S ? kc[s] : 0ull;
    // (a thread block never straddles two permutation blocks: perm is a multiple of 256)
    const int b0 = (blockIdx.x * blockDim.x) / perm * perm, b1 = min(b0 + perm, S);
    int rank = 0;
    for (int j0 = b0; j0 < b1; j0 += 256) {
        __syncthreads();
        s_key[threadIdx.x] = j0 + (int)threadIdx.x < b1 ? kc[j0 + threadIdx.x] : ~0ull;
        __syncthreads();
        const int nj = min(256, b1 - j0);
        for (int j = 0; j < nj; ++j) rank += s_key[j] < mine;          // keys are distinct (the draw index is part of them)
    }
    if (s < S) {
        draw_slot[(size_t)c * S + s] = b0 + rank;
        slot_draw[(size_t)c * n_slots + b0 + rank] = s;
    }
}

// one warp per (class, slice): the slice's row count = the longest of its 32 draws (class D: also its deepest node)
__global__ void __launch_bounds__(128) bs_hdr_kernel(const BsCnt* __restrict__ draw_tot, const int32_t* __restrict__ slot_draw,
                                                     int n_slices, int2* __restrict__ slice_hdr) {
    const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= BS_NCLS * n_slices) return;
    const int c = w / n_slices;
    const int s = slot_draw[(size_t)w * 32 + lane];           // [class][slice][lane]
    int rows = 0, dmax = 0;
    if (s >= 0) {
        const BsCnt t = draw_tot[s];
        rows = t.n[c];
        dmax = c == BS_D ? t.dmax : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0) slice_hdr[w] = make_int2(rows, dmax);
}

__global__ void __launch_bounds__(128) bs_fill_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                      const int32_t* __restrict__ tree_leaf_off, const double* __restrict__ leaf_mu,
                                                      const int32_t* __restrict__ cut_off, int64_t n_trees, int m, int S,
                                                      const BsCnt* __restrict__ tree_row, const int32_t* __restrict__ draw_slot,
                                                      BsLayout L, double fx_scale, unsigned char* __restrict__ pool) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_trees) return;
    const int s_draw = (int)(k / m);
    int slot[BS_NCLS];
#pragma unroll
    for (int q = 0; q < BS_NCLS; ++q) slot[q] = draw_slot[(size_t)q * S + s_draw];
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
                         uint32_t* __restrict__ anc = reinterpret_cast<uint32_t*>(pool + L.anc_off) +
                                                      (int64_t)(slot[BS_D] >> 5) * L.anc_planes * L.cap[BS_D] * 32 + (slot[BS_D] & 31);
                         for (int a = 0; a < L.anc_planes; ++a)
                             anc[((int64_t)a * L.cap[BS_D] + row[BS_D]) * 32] = a < depth ? path[a] : (uint32_t)L.ones_off;
                     }
                     int rr = 0, sl = 0, cap = 1;
                     int64_t off = 0;
#pragma unroll
                     for (int q = 0; q < BS_NCLS; ++q)
                         if (q == c) { rr = row[q]; row[q] = rr + 1; sl = slot[q]; cap = L.cap[q]; off = L.cls_off[q]; }
                     reinterpret_cast<uint4*>(pool + off)[((int64_t)(sl >> 5) * cap + rr) * 32 + (sl & 31)] = e;
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
    int ch = 0, perm = 0;
    for (int c : {8, 4, 2}) {                          // the widest tile that fits, then the largest permutation block
        for (int p : {BS_PERM, BS_PERM_SMALL})
            if ((size_t)(ncuts + 1) * (size_t)(c * 16 + 4) + 64 + (size_t)f->d * (size_t)c * 128 + (size_t)p * 26 + 2048 <= BS_SMEM_MAX) {
                ch = c; perm = p; break;
            }
        if (ch) break;
    }
    if (ch == 0 || (ncuts + 1) * (int64_t)ch >= 1 << 14) return;         // row offsets are 14-bit counts of 16 bytes
    f->bs_ncuts = (int32_t)ncuts;
    f->bs_ch = ch;
    f->bs_perm = perm;
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
    const int n_slices = L.n_slices, n_slots = n_slices * 32;
    const size_t pool_bytes = (size_t)L.total_bytes;
    BsCnt *tree_cnt = nullptr, *tree_row = nullptr, *draw_tot = nullptr;
    unsigned long long* key = nullptr;
    int32_t* draw_slot = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_pool, pool_bytes, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_slice_hdr, sizeof(int2) * (size_t)BS_NCLS * (size_t)n_slices, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_base, sizeof(double) * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&f->dev.bs_slot_draw, sizeof(int32_t) * (size_t)BS_NCLS * (size_t)n_slots, st));
    BI_CUDA(cudaMallocAsync((void**)&tree_cnt, sizeof(BsCnt) * (size_t)n_trees, st));
    BI_CUDA(cudaMallocAsync((void**)&tree_row, sizeof(BsCnt) * (size_t)n_trees, st));
    BI_CUDA(cudaMallocAsync((void**)&draw_tot, sizeof(BsCnt) * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&key, sizeof(unsigned long long) * (size_t)BS_NCLS * (size_t)f->S, st));
    BI_CUDA(cudaMallocAsync((void**)&draw_slot, sizeof(int32_t) * (size_t)BS_NCLS * (size_t)f->S, st));
    BI_CUDA(cudaMemsetAsync(f->dev.bs_pool, 0, pool_bytes, st));
    BI_CUDA(cudaMemsetAsync(f->dev.bs_slot_draw, 0xFF, sizeof(int32_t) * (size_t)BS_NCLS * (size_t)n_slots, st));   // -1: empty slot
    const unsigned gt = (unsigned)((n_trees + 127) / 128);
    bs_count_kernel<<<gt, 128, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.cut_off, n_trees, tree_cnt);
    bs_draw_kernel<<<(unsigned)((f->S + 127) / 128), 128, 0, st>>>(tree_cnt, f->dev.tree_leaf_off, f->dev.leaf_mu, f->S, f->m,
                                                                  tree_row, draw_tot, key, f->dev.bs_base);
    bs_rank_kernel<<<dim3((unsigned)((f->S + 255) / 256), BS_NCLS), 256, 0, st>>>(key, f->S, draw_slot, f->dev.bs_slot_draw, n_slots, f->bs_perm);
    bs_hdr_kernel<<<(unsigned)((BS_NCLS * n_slices * 32 + 127) / 128), 128, 0, st>>>(draw_tot, f->dev.bs_slot_draw, n_slices,
                                                                                      f->dev.bs_slice_hdr);
    bs_fill_kernel<<<gt, 128, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off, f->dev.leaf_mu, f->dev.cut_off,
                                       n_trees, f->m, f->S, tree_row, draw_slot, L, std::ldexp(1.0, f->bs_q), f->dev.bs_pool);
    count_launch(5);
    for (void* q : {(void*)tree_cnt, (void*)tree_row, (void*)draw_tot, (void*)key, (void*)draw_slot}) cudaFreeAsync(q, st);
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// what the slices load, in mask-row references (for bartint_forest_bitslice_info): every class padded to the longest
// draw of each of its slices.  Reads the slice headers back (blocks).
int64_t bs_padded_refs(const bartint_forest* f) {
    if (!f->bs_ok || f->dev.bs_slice_hdr == nullptr) return 0;
    const int n_slices = (f->S + 31) / 32;
    std::vector<int2> hdr((size_t)BS_NCLS * (size_t)n_slices);
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(hdr.data(), f->dev.bs_slice_hdr, hdr.size() * sizeof(int2), cudaMemcpyDeviceToHost) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int64_t padded = 0;
    const int refs[BS_NCLS] = {0, 2, 3, 4, 1};
    for (int c = 1; c < BS_NCLS; ++c)
        for (int sl = 0; sl < n_slices; ++sl) {
            const int2 h = hdr[(size_t)c * n_slices + sl];
            padded += 32ll * h.x * (refs[c] + (c == BS_D ? h.y : 0));
        }
    return padded;
}

// ---------------------------------------------------------------------------------------------
// the evaluation kernel
// ---------------------------------------------------------------------------------------------
struct BsParams {
    const unsigned char* pool;
    const int2* slice_hdr;      // [class][slice]: rows, deepest node (class D)
    const int32_t* slot_draw;   // [class][slice * 32 + lane] -> draw or -1
    BsLayout L;
    const uint4* codes_pm;      // point-major packed codes (gather forests), or null
    const uint8_t* codes;       // feature-major code rows
    int64_t Npad, N;
    const int32_t* cut_off;
    int d, ncuts;
    int draw_begin, draw_end;
    int n_slices, unit_chunks, perm;  // grid.y = permutation blocks x unit_chunks; perm = draws per block
    double inv_scale;           // 2^-Q
    double* draw_part;          // [n_tiles][draw_end - draw_begin]
};

// shared-memory layout of the evaluation kernel: [mask rows][row counts][slice counter][staged codes d x TILE]
__host__ __device__ inline size_t bs_stage_off(int n_rows, int rowb) { return (((size_t)n_rows * (size_t)(rowb + 4) + 32 + 15) & ~(size_t)15); }

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
// 16 bytes at offset a ^ x of the CTA's shared memory.  A lane reads the chunks of a mask row in the order k ^ (lane % CH): the 8
// lanes of an LDS.128 phase then hit 8 different bank groups.  Rows are aligned to their own size, so with
// a = row + 16 * (lane % CH) chunk k ^ (lane % CH) is at a ^ (16 k): one LOP3 with an immediate, no table of offsets.
#define lds16x(a, x) (*reinterpret_cast<const uint4*>(smem + ((a) ^ (uint32_t)(x))))

// Rows [0, n) of one class, G at a time: the loads of the next group are in flight while this one is processed (the
// row blocks are padded to a multiple of G with zero entries, so a group may always be loaded whole).
template <int G, typename F>
__device__ __forceinline__ void bs_rows(const uint4* __restrict__ rows, int n, F body) {
    uint4 w[G];
    if (n > 0) {
#pragma unroll
        for (int j = 0; j < G; ++j) w[j] = __ldg(rows + (size_t)j * 32);
    }
    for (int k0 = 0; k0 < n; k0 += G) {
        const bool more = k0 + G < n;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint4 e = w[j];                  // entry k0 + j; its register slot takes entry k0 + G + j at once
            if (more) w[j] = __ldg(rows + (size_t)(k0 + G + j) * 32);
            if (k0 + j < n) body(e);
        }
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
    unsigned long long* const s_acc =                                                   // [perm][2] integer accumulators
        reinterpret_cast<unsigned long long*>(smem + ((bs_stage_off(n_rows, ROWB) + (size_t)prm.d * TILE + 15) & ~(size_t)15));
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
    const uint32_t mb = (uint32_t)(lane & (CH - 1)) * 16u;          // (the mask table starts at offset 0, 128-byte aligned)
    // ---- 3. per-row popcounts over the tile (class A reads these instead of the masks) ----
    for (int r = tid; r < n_rows; r += BS_THREADS) {
        const uint32_t row = mb + (uint32_t)r * ROWB;
        int cnt = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) cnt += popc4(lds16x(row, c * 16));
        s_cnt[r] = (uint32_t)cnt;
    }
    __syncthreads();

    // ---- 4. the rows.  A WORK UNIT is one (class, slice): 32 draws with (nearly) the same number of rows of that class,
    // one lane per draw.  The warps take the units of this CTA's permutation block from a shared counter -- the classes
    // in the order B, C2, C3, D, A and the slices of a class from the longest down -- and add the lane's exact integer
    // sums to its draw's accumulators in shared memory (64-bit integer atomics: the order of arrival does not matter).
    // Each class sorts the draws on its own, so a slice pads only to the longest of 32 nearly equal draws; with one
    // common assignment of draws to slices the classes padded 7.8e7 wavefronts for 5.9e7 needed at cfg 2. ----
    // CTA constants of the unit loop live in shared memory and are re-read where needed: the loop bodies need every
    // register for mask words in flight (128 per thread at 512 threads)
    volatile int* const s_ctl = s_next + 1;             // [0] first slice, [1] slices, [2] units, [3] unit chunk
    if (tid == 0) {
        const int pblock = prm.draw_begin / prm.perm + blockIdx.y / prm.unit_chunks;
        const int a = pblock * (prm.perm / 32), b = min(a + prm.perm / 32, prm.n_slices);
        s_ctl[0] = a; s_ctl[1] = b - a; s_ctl[2] = BS_NCLS * (b - a); s_ctl[3] = blockIdx.y % prm.unit_chunks;
    }
    __syncthreads();
    {
        const int sl0 = s_ctl[0], nsl = s_ctl[1], draw0 = sl0 * 32;
        for (int i = tid; i < nsl * 64; i += BS_THREADS) s_acc[i] = 0ull;
        // the units' slice headers and slot -> draw maps, staged once: a unit would otherwise start with two dependent L2
        // round trips, and a CTA runs up to 5 x 32 of them
        int2* const s_hdr = reinterpret_cast<int2*>(s_acc + prm.perm * 2);                  // [class][slice of the block]
        uint16_t* const s_slot = reinterpret_cast<uint16_t*>(s_hdr + BS_NCLS * (prm.perm / 32));   // [class][slot] -> draw - draw0
        for (int i = tid; i < BS_NCLS * nsl; i += BS_THREADS)
            s_hdr[i] = __ldg(&prm.slice_hdr[(size_t)(i / nsl) * prm.n_slices + sl0 + i % nsl]);
        for (int i = tid; i < BS_NCLS * nsl * 32; i += BS_THREADS) {
            const int sd = __ldg(&prm.slot_draw[(size_t)(i / (nsl * 32)) * prm.n_slices * 32 + draw0 + i % (nsl * 32)]);
            s_slot[i] = sd < 0 ? (uint16_t)0xFFFFu : (uint16_t)(sd - draw0);
        }
    }
    __syncthreads();
    for (;;) {
        int take = 0;
        if (lane == 0) take = atomicAdd(s_next, 1);
        take = __shfl_sync(0xffffffffu, take, 0);
        int u = take * prm.unit_chunks + s_ctl[3];                 // this CTA's share of the units
        if (u >= s_ctl[2]) break;
        // units in the order (slice rank descending; B, C2, C3, D, A of that rank): at any moment the warps work on a
        // mix of classes -- class A uses no mask rows, and a CTA whose warps all run A units leaves the data pipe idle
        const int ci = u % BS_NCLS;
        const int c = ci == 4 ? BS_A : ci + 1;
        int2 hd;
        int sl;
        {
            const int nsl = s_ctl[1];
            const int idx = c * nsl + nsl - 1 - u / BS_NCLS;
            const int2* s_hdr = reinterpret_cast<const int2*>(s_acc + prm.perm * 2);
            const uint16_t* s_slot = reinterpret_cast<const uint16_t*>(s_hdr + BS_NCLS * (prm.perm / 32));
            const int so = s_slot[idx * 32 + lane];
            sl = s_ctl[0] + nsl - 1 - u / BS_NCLS;
            const int s = so == 0xFFFF ? -1 : s_ctl[0] * 32 + so;
            if (!__any_sync(0xffffffffu, s >= prm.draw_begin && s < prm.draw_end)) continue;
            hd = s_hdr[idx];
        }
        if (hd.x == 0) continue;
        const uint4* __restrict__ rows =
            reinterpret_cast<const uint4*>(prm.pool + prm.L.cls_off[c]) + (size_t)sl * prm.L.cap[c] * 32 + lane;
        unsigned long long acc0 = 0ull;      // sum of (low 32 bits of delta) * count
        long long acc1 = 0ll;                // sum of (high 32 bits of delta, signed) * count
        auto add = [&](const uint4& e, int cnt) {
            acc0 += (unsigned long long)e.x * (uint32_t)cnt;
            acc1 += (long long)(int)e.y * (long long)cnt;
        };
        if (c == BS_A) {
            bs_rows<8>(rows, hd.x, [&](const uint4& e) { add(e, (int)s_cnt[e.z]); });
        } else if (c == BS_B) {
            bs_rows<2>(rows, hd.x, [&](const uint4& e) {
                const uint32_t pm = (uint32_t)((int)e.z >> 31);
                const uint32_t ap = mb + (e.z & 0x7FFFFFFFu), aj = mb + e.w;
                PopAcc pa;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const uint4 mp = lds16x(ap, k * 16), mj = lds16x(aj, k * 16);
                    pa.add4((mp.x ^ pm) & mj.x, (mp.y ^ pm) & mj.y, (mp.z ^ pm) & mj.z, (mp.w ^ pm) & mj.w);
                }
                add(e, pa.total());
            });
        } else if (c == BS_C2) {
            bs_rows<4>(rows, hd.x, [&](const uint4& e) {
                const uint32_t aj = mb + ((e.z & 0x3FFFu) << 4), a0 = mb + (((e.z >> 14) & 0x3FFFu) << 4),
                               a1 = mb + ((e.w & 0x3FFFu) << 4);
                const uint32_t p0 = (uint32_t)((int)(e.z << 3) >> 31), p1 = (uint32_t)((int)(e.z << 2) >> 31);
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const uint4 mj = lds16x(aj, k * 16), m0 = lds16x(a0, k * 16), m1 = lds16x(a1, k * 16);
                    cnt += __popc((m0.x ^ p0) & (m1.x ^ p1) & mj.x) + __popc((m0.y ^ p0) & (m1.y ^ p1) & mj.y) +
                           __popc((m0.z ^ p0) & (m1.z ^ p1) & mj.z) + __popc((m0.w ^ p0) & (m1.w ^ p1) & mj.w);
                }
                add(e, cnt);
            });
        } else if (c == BS_C3) {
            bs_rows<2>(rows, hd.x, [&](const uint4& e) {
                const uint32_t aj = mb + ((e.z & 0x3FFFu) << 4), a0 = mb + (((e.z >> 14) & 0x3FFFu) << 4),
                               a1 = mb + ((e.w & 0x3FFFu) << 4), a2 = mb + (((e.w >> 14) & 0x3FFFu) << 4);
                const uint32_t p0 = (uint32_t)((int)(e.z << 3) >> 31), p1 = (uint32_t)((int)(e.z << 2) >> 31),
                               p2 = (uint32_t)((int)(e.z << 1) >> 31);
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const uint4 mj = lds16x(aj, k * 16), m0 = lds16x(a0, k * 16), m1 = lds16x(a1, k * 16), m2 = lds16x(a2, k * 16);
                    cnt += __popc((m0.x ^ p0) & (m1.x ^ p1) & (m2.x ^ p2) & mj.x) + __popc((m0.y ^ p0) & (m1.y ^ p1) & (m2.y ^ p2) & mj.y) +
                           __popc((m0.z ^ p0) & (m1.z ^ p1) & (m2.z ^ p2) & mj.z) + __popc((m0.w ^ p0) & (m1.w ^ p1) & (m2.w ^ p2) & mj.w);
                }
                add(e, cnt);
            });
        } else {      // class D: depth >= 4, ancestors in planes (rare)
            const uint32_t* __restrict__ anc = reinterpret_cast<const uint32_t*>(prm.pool + prm.L.anc_off) +
                                               (size_t)sl * prm.L.anc_planes * prm.L.cap[BS_D] * 32 + lane;
            for (int k = 0; k < hd.x; ++k) {
                const uint4 e = __ldg(rows + (size_t)k * 32);
                uint4 r[CH];
                const uint32_t aj = mb + e.z;
#pragma unroll
                for (int q = 0; q < CH; ++q) r[q] = lds16x(aj, q * 16);
                for (int a = 0; a < hd.y; ++a) {
                    const uint32_t w = __ldg(anc + ((size_t)a * prm.L.cap[BS_D] + k) * 32);
                    const uint32_t pm = (uint32_t)((int)w >> 31);
                    const uint32_t ap = mb + (w & 0x7FFFFFFFu);
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const uint4 mp = lds16x(ap, q * 16);
                        r[q].x &= mp.x ^ pm; r[q].y &= mp.y ^ pm; r[q].z &= mp.z ^ pm; r[q].w &= mp.w ^ pm;
                    }
                }
                int cnt = 0;
#pragma unroll
                for (int q = 0; q < CH; ++q) cnt += popc4(r[q]);
                add(e, cnt);
            }
        }
        // the unit's draw slot again (nothing derived from u was kept in registers across the rows)
        asm volatile("" : "+r"(u));
        {
            const int nsl = s_ctl[1], cj = u % BS_NCLS;
            const int idx = (cj == 4 ? BS_A : cj + 1) * nsl + nsl - 1 - u / BS_NCLS;
            const uint16_t* s_slot = reinterpret_cast<const uint16_t*>(s_acc + prm.perm * 2 + BS_NCLS * (prm.perm / 32));
            const int so = s_slot[idx * 32 + lane];
            if (so != 0xFFFF) {
                atomicAdd(&s_acc[so * 2], acc0);
                atomicAdd(&s_acc[so * 2 + 1], (unsigned long long)acc1);
            }
        }
    }
    __syncthreads();
    {
        const int nsl = s_ctl[1], draw0 = s_ctl[0] * 32, uchunk = s_ctl[3];
        const int ndraws = prm.draw_end - prm.draw_begin;
        for (int i = threadIdx.x; i < nsl * 32; i += BS_THREADS) {
            const int s = draw0 + i;
            if (s >= prm.draw_begin && s < prm.draw_end)
                prm.draw_part[((int64_t)blockIdx.x * prm.unit_chunks + uchunk) * ndraws + (s - prm.draw_begin)] =
                    ((double)(long long)s_acc[2 * i + 1] * 4294967296.0 + (double)s_acc[2 * i]) * prm.inv_scale;
        }
    }
}

size_t bs_smem_bytes(const bartint_forest* f) {
    return ((bs_stage_off(f->bs_ncuts + 1, f->bs_ch * 16) + (size_t)f->d * (size_t)f->bs_ch * 128 + 15) & ~(size_t)15) +
           (size_t)f->bs_perm * 16 + (size_t)BS_NCLS * (f->bs_perm / 32) * 8 + (size_t)BS_NCLS * f->bs_perm * 2;
}
int bs_tile(const bartint_forest* f) { return f->bs_ch * 128; }
int64_t bs_pool_bytes(const bartint_forest* f) { return f->bs_ok ? bs_layout_of(f).total_bytes : 0; }

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
    prm.perm = f->bs_perm;
    prm.unit_chunks = plan.chunk_draws;           // (unit chunks per permutation block for this kernel)
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
__global__ void __launch_bounds__(1024) bs_reduce_kernel(const double* __restrict__ part, int n_tiles, int ndraws,
                                                         const double* __restrict__ base, double n_points,
                                                         double* __restrict__ out) {
    // 8 draws x 128 tile lanes per CTA: a thread adds every 128th tile of its draw (8 loads for 1024 tiles, all in
    // flight at once; 8 lanes read one 64-byte piece of a row), then the partial sums are combined in a fixed order
    __shared__ double s_part[128][9];
    __shared__ double s_mid[8][9];
    const int x = threadIdx.x & 7, y = threadIdx.x >> 3;
    const int s = blockIdx.x * 8 + x;
    const int t0 = blockIdx.y * BS_RED_BLOCK, t1 = min(t0 + BS_RED_BLOCK, n_tiles);
    double acc = 0.0;
    if (s < ndraws) {
#pragma unroll 8
        for (int t = t0 + y; t < t1; t += 128) acc += part[(int64_t)t * ndraws + s];
    }
    s_part[y][x] = acc;
    __syncthreads();
    if (threadIdx.x < 64) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) tot += s_part[y * 16 + k][x];
        s_mid[y][x] = tot;
    }
    __syncthreads();
    if (threadIdx.x < 8 && s < ndraws) {
        double tot = s_mid[0][x];
#pragma unroll
        for (int k = 1; k < 8; ++k) tot += s_mid[k][x];
        out[(int64_t)blockIdx.y * ndraws + s] = base != nullptr ? tot + n_points * base[s] : tot;
    }
}

int launch_bs_reduce(const bartint_forest* f, const double* draw_part, int n_tiles, int32_t draw_begin, int ndraws,
                     int64_t N, double* draw_sum, cudaStream_t st) {
    if (ndraws == 0) return BARTINT_OK;
    const double* base = f->dev.bs_base + draw_begin;
    const int n_blocks = (n_tiles + BS_RED_BLOCK - 1) / BS_RED_BLOCK;
    if (n_blocks <= 1) {
        bs_reduce_kernel<<<(ndraws + 7) / 8, 1024, 0, st>>>(draw_part, n_tiles, ndraws, base, (double)N, draw_sum);
        count_launch();
    } else {
        BI_REQUIRE(n_blocks <= BS_RED_BLOCK, BARTINT_ERR_UNSUPPORTED, "more than %d tiles in one evaluation", BS_RED_BLOCK * BS_RED_BLOCK);
        double* tmp = nullptr;
        BI_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * (size_t)n_blocks * (size_t)ndraws, st));
        bs_reduce_kernel<<<dim3((unsigned)((ndraws + 7) / 8), (unsigned)n_blocks), 1024, 0, st>>>(draw_part, n_tiles, ndraws, nullptr,
                                                                                                   0.0, tmp);
        bs_reduce_kernel<<<(ndraws + 7) / 8, 1024, 0, st>>>(tmp, n_blocks, ndraws, base, (double)N, draw_sum);
        count_launch(2);
        cudaFreeAsync(tmp, st);
    }
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
