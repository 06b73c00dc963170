// K0 point binning, table builder, K2 exact leaf-probability integral, K3 merges/finalisation.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace bartint {

__device__ __forceinline__ double warp_sum_misc(double v) {        // fixed-order butterfly: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// K0  bin_points: code = #{k : x > cut[j][k]}  (dbarts' integer cut map of x.test).
// HBM-bound streaming kernel: reads 8 B, writes 1 B per matrix element.  One thread = 4 consecutive
// points of one variable -> one 32-bit word of the feature-major code array (the word the table
// kernel's byte-SIMD compare consumes).  Grid: x over point quads, y over variables.
// ---------------------------------------------------------------------------------------------
// Guess-and-verify: the answer is ALWAYS decided by comparisons against the real cut values (exact for any ascending
// cut list); the affine guess only makes the common case two probes on (near-)uniform grids such as dbarts'
// usequants=FALSE cut points, instead of the 7 dependent probes of a binary search (which made the kernel
// instruction/shared-memory bound rather than HBM bound).  `cuts` is sentinel padded: cuts[-1] = -inf, cuts[nc] = +inf,
// so  k is correct  <=>  cuts[k-1] < x  &&  !(cuts[k] < x).
// INC = false (dbarts):      code = #{k : cut[k] <  x}, NaN -> 0   ("right iff x >  cut"  <=>  code > cut index)
// INC = true  (BART::wbart):  code = #{k : cut[k] <= x}, NaN -> nc  ("right iff x >= cut", tree::bn's "x < c -> left")
template <bool INC>
__device__ __forceinline__ bool cut_below(double c, double x) { return INC ? c <= x : c < x; }

template <bool INC>
__device__ __forceinline__ uint32_t bin_one(double x, const double* __restrict__ cuts, int nc, double c0, double inv_step) {
    if (INC && x != x) return (uint32_t)nc;               // NaN: every "x < c" is false -> always right
    const double g = fma(x - c0, inv_step, 1.0);          // uniform grid: #{k : cut[k] < x} = floor((x - c0)/step) + 1
    int k = __double2int_rd(fmin(fmax(g, 0.0), (double)nc));   // NaN -> fmax/fmin return the non-NaN bound -> 0
    const bool up = cut_below<INC>(cuts[k], x), lo_ok = cut_below<INC>(cuts[k - 1], x);
    if (lo_ok && !up) return (uint32_t)k;                 // common case: guess was right
    if (up && !cut_below<INC>(cuts[k + 1 > nc ? nc : k + 1], x)) return (uint32_t)(k + 1);
    if (!lo_ok && k >= 2 && cut_below<INC>(cuts[k - 2], x)) return (uint32_t)(k - 1);
    int lo = 0, len = nc;                                 // irregular grid / NaN / infinities: plain lower-bound search
    while (len > 0) {
        int h = len >> 1;
        bool r = cut_below<INC>(cuts[lo + h], x);
        lo = r ? lo + h + 1 : lo;
        len = r ? len - h - 1 : h;
    }
    return (uint32_t)lo;
}

// Single-precision screen in front of bin_one: pair[k] = {float_ru(cuts[k-1]), float_rd(cuts[k])}, one 8-byte probe
// instead of two, and the guess itself in fp32.  With xn = float_rn(x):  |x - xn| <= ulp/2, so x lies strictly between
// the float neighbours of xn, and for floats U < xn < D:
//     U <= prev(xn) < x   =>  cuts[k-1] <= U < x          x < next(xn) <= D  =>  x < D <= cuts[k]
// i.e. a passed screen PROVES cuts[k-1] < x < cuts[k], which decides the code under both routing rules.  Anything
// else (x within a float ulp of a cut or exactly on it, wrong guess, NaN, +-Inf, irregular grid) takes the exact fp64
// path.  ncu before the screen: cut-table probes at 70 % of the L1/shared pipe with 55 % bank-conflict replays,
// ahead of HBM; with two-sided fp64 guess arithmetic the kernel was then issue-bound (78 %), hence the fp32 guess.
template <bool INC>
__device__ __forceinline__ uint32_t bin_screened(double x, const float2* __restrict__ pair, const double* __restrict__ cuts,
                                                 int nc, double c0, double inv_step, float c0f, float invf) {
    const float xn = __double2float_rn(x);
    const int k = min(max(__float2int_rd(fmaf(xn - c0f, invf, 1.0f)), 0), nc);     // NaN -> 0
    const float2 pr = pair[k];
    if (pr.x < xn && xn < pr.y) return (uint32_t)k;
    return bin_one<INC>(x, cuts, nc, c0, inv_step);
}

// fills the sentinel-padded fp64 cut row and the float pair table of one variable (all threads of the CTA)
__device__ __forceinline__ void load_cut_tables(const double* __restrict__ cut_val, int c0i, int nc, double* s_cut,
                                                float2* s_pair) {
    for (int k = threadIdx.x; k < nc; k += blockDim.x) s_cut[k] = cut_val[c0i + k];
    if (threadIdx.x == 0) { s_cut[-1] = -INFINITY; s_cut[nc] = INFINITY; }
    for (int k = threadIdx.x; k <= nc; k += blockDim.x) {
        const double below = k > 0 ? cut_val[c0i + k - 1] : -INFINITY, here = k < nc ? cut_val[c0i + k] : INFINITY;
        s_pair[k] = make_float2(__double2float_ru(below), __double2float_rd(here));
    }
}

template <bool INC>
__global__ void __launch_bounds__(256) bin_points_kernel(const double* __restrict__ X, int64_t ld, int64_t N,
                                                         const int32_t* __restrict__ cut_off,
                                                         const double* __restrict__ cut_val,
                                                         uint32_t* __restrict__ codes4, int64_t quads_per_row,
                                                         int64_t n_quads) {
    __shared__ double s_cut_pad[258];
    __shared__ float2 s_pair[257];
    double* s_cut = s_cut_pad + 1;
    const int j = blockIdx.y;
    const int c0i = cut_off[j], nc = cut_off[j + 1] - c0i;
    load_cut_tables(cut_val, c0i, nc, s_cut, s_pair);
    __syncthreads();
    const double c0 = nc > 0 ? s_cut[0] : 0.0;
    const double span = nc > 1 ? s_cut[nc - 1] - s_cut[0] : 0.0;
    const double inv_step = span > 0.0 ? (double)(nc - 1) / span : 0.0;
    const float c0f = (float)c0, invf = (float)inv_step;
    const double* __restrict__ col = X + (int64_t)j * ld;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(col) & 15) == 0);
    // quads_per_row = stride of the code rows; n_quads = words this launch writes (a row block of a larger point set)
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_quads;
         w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = w * 4;
        uint32_t word = 0;
        if (i + 3 < N) {
            double x0, x1, x2, x3;
            if (aligned16) {
                double2 a = __ldg(reinterpret_cast<const double2*>(col + i));
                double2 b = __ldg(reinterpret_cast<const double2*>(col + i + 2));
                x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y;
            } else {
                x0 = __ldg(col + i); x1 = __ldg(col + i + 1); x2 = __ldg(col + i + 2); x3 = __ldg(col + i + 3);
            }
            word = bin_screened<INC>(x0, s_pair, s_cut, nc, c0, inv_step, c0f, invf) |
                   (bin_screened<INC>(x1, s_pair, s_cut, nc, c0, inv_step, c0f, invf) << 8) |
                   (bin_screened<INC>(x2, s_pair, s_cut, nc, c0, inv_step, c0f, invf) << 16) |
                   (bin_screened<INC>(x3, s_pair, s_cut, nc, c0, inv_step, c0f, invf) << 24);
        } else {
            for (int k = 0; k < 4; ++k)
                if (i + k < N) word |= bin_screened<INC>(__ldg(col + i + k), s_pair, s_cut, nc, c0, inv_step, c0f, invf) << (8 * k);
        }
        codes4[(int64_t)j * quads_per_row + w] = word;   // padding quads (i >= N) are written as 0
    }
}

int launch_bin_points(const double* dX, int64_t ld, int64_t N, int32_t d, const int32_t* d_cut_off,
                      const double* d_cut_val, int max_cuts, int route, uint8_t* codes, int64_t Npad, cudaStream_t st,
                      int64_t rows_padded) {
    (void)max_cuts;
    if (d == 0 || Npad == 0) return BARTINT_OK;
    // rows_padded > 0: bin a row block -- `codes` points at the block's first row inside code rows of stride Npad, and
    // rows_padded (a multiple of 4, >= N) rows are written (zeros beyond N)
    const int64_t quads = (rows_padded > 0 ? rows_padded : Npad) / 4;
    int64_t bx = (quads + 255) / 256;
    // a few waves of 148 SMs x 8 resident CTAs; grid-stride beyond that
    const int64_t cap = (148 * 8 * 4 + d - 1) / d;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)d);
    auto kern = route == BARTINT_ROUTE_BART ? bin_points_kernel<true> : bin_points_kernel<false>;
    kern<<<grid, 256, 0, st>>>(dX, ld, N, d_cut_off, d_cut_val, reinterpret_cast<uint32_t*>(codes), Npad / 4, quads);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// K0 for gather forests (d <= 16): one pass from the fp64 matrix straight to the point-major packed records the gather
// kernel keeps in registers (8 B read per element, 16 B written per point); the feature-major code rows are only
// materialised on demand (unpack_codes_kernel: walk kernel, leaf decoding, bartint_points_codes).  Persistent grid,
// ONE point per thread and iteration -- fine-grained on purpose: with fatter threads (4 points x all variables) the
// kernel is a couple of waves of long CTAs and loses a third of its time to the last, partly filled wave.  All d cut
// lists sit in shared memory twice: float pairs for the screen and sentinel-padded fp64 rows for the exact path.
// ---------------------------------------------------------------------------------------------
template <bool INC, int DMAX>
__global__ void __launch_bounds__(256) bin_pack_points_kernel(const double* __restrict__ X, int64_t ld, int64_t N, int d,
                                                              const int32_t* __restrict__ cut_off,
                                                              const double* __restrict__ cut_val, int cut_stride,
                                                              uint4* __restrict__ codes_pm, int64_t rows) {
    extern __shared__ double s_dyn[];                 // d fp64 rows [cut_stride], then d float2 rows [cut_stride]
    __shared__ double s_c0[16], s_inv[16];            // exact path only
    __shared__ float4 s_fc[16];                       // per variable: {c0, 1/step, nc (int bits), -} for the fp32 guess
    float2* s_pairs = reinterpret_cast<float2*>(s_dyn + (size_t)d * cut_stride);
    // prologue: one warp per variable (8 in flight), so the dependent loads cut_off -> cut_val of the d variables overlap
    // instead of costing d round trips in front of every CTA (measured: 20 us of a 39 us kernel at 10^6 points)
    {
        const int lane = threadIdx.x & 31;
        for (int j = threadIdx.x >> 5; j < d; j += blockDim.x >> 5) {
            const int c0i = cut_off[j], nc = cut_off[j + 1] - c0i;
            double* row = s_dyn + (size_t)j * cut_stride + 1;
            float2* pr = s_pairs + (size_t)j * cut_stride;
            for (int k = lane; k < nc; k += 32) row[k] = cut_val[c0i + k];
            for (int k = lane; k <= nc; k += 32) {
                const double below = k > 0 ? cut_val[c0i + k - 1] : -INFINITY, here = k < nc ? cut_val[c0i + k] : INFINITY;
                pr[k] = make_float2(__double2float_ru(below), __double2float_rd(here));
            }
            if (lane == 0) {
                row[-1] = -INFINITY;
                row[nc] = INFINITY;
                const double first = nc > 0 ? cut_val[c0i] : 0.0;
                const double span = nc > 1 ? cut_val[c0i + nc - 1] - first : 0.0;
                const double inv = span > 0.0 ? (double)(nc - 1) / span : 0.0;
                s_c0[j] = first;
                s_inv[j] = inv;
                s_fc[j] = make_float4((float)first, (float)inv, __int_as_float(nc), 0.f);
            }
        }
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (i < N) {
            double x[DMAX];
#pragma unroll
            for (int j = 0; j < DMAX; ++j)                            // all loads first: d x 8 B in flight per thread
                if (j < d) x[j] = __ldg(X + (int64_t)j * ld + i);
#pragma unroll
            for (int j = 0; j < DMAX; ++j)
                if (j < d) {
                    // the screen of bin_screened(), with the per-variable constants as one broadcast 16-byte load
                    const float4 fc = s_fc[j];
                    const int nc = __float_as_int(fc.z);
                    const float xn = __double2float_rn(x[j]);
                    const int k = min(max(__float2int_rd(fmaf(xn - fc.x, fc.y, 1.0f)), 0), nc);
                    const float2 pr = s_pairs[j * cut_stride + k];
                    uint32_t code = (uint32_t)k;
                    if (!(pr.x < xn && xn < pr.y))
                        code = bin_one<INC>(x[j], s_dyn + (size_t)j * cut_stride + 1, nc, s_c0[j], s_inv[j]);
                    w[j >> 2] |= code << (8 * (j & 3));
                }
        }
        codes_pm[i] = make_uint4(w[0], w[1], w[2], w[3]);             // rows beyond N: zero records
    }
}

int launch_bin_pack_points(const double* dX, int64_t ld, int64_t N, int32_t d, const int32_t* d_cut_off,
                           const double* d_cut_val, int max_cuts, int route, uint4* codes_pm, int64_t rows,
                           cudaStream_t st) {
    if (rows <= 0) return BARTINT_OK;
    BI_REQUIRE(d >= 1 && d <= 16, BARTINT_ERR_ARG, "internal: packed binning needs 1 <= d <= 16");
    const int cut_stride = max_cuts + 2;
    const size_t smem = (size_t)d * cut_stride * (sizeof(double) + sizeof(float2));
    typedef void (*kern_t)(const double*, int64_t, int64_t, int, const int32_t*, const double*, int, uint4*, int64_t);
    const bool inc = route == BARTINT_ROUTE_BART;
    kern_t k = d <= 4    ? (inc ? bin_pack_points_kernel<true, 4> : bin_pack_points_kernel<false, 4>)
               : d <= 8  ? (inc ? bin_pack_points_kernel<true, 8> : bin_pack_points_kernel<false, 8>)
               : d <= 12 ? (inc ? bin_pack_points_kernel<true, 12> : bin_pack_points_kernel<false, 12>)
                         : (inc ? bin_pack_points_kernel<true, 16> : bin_pack_points_kernel<false, 16>);
    // occupancy and SM count are cached per (kernel, shared-memory size): the queries cost ~15 us of host time, which a
    // 25 us kernel cannot hide
    struct Cached { kern_t k = nullptr; size_t smem = 0; int dev = -1, per_sm = 1, n_sm = 148; };
    static thread_local Cached cache;
    int dev_now = 0;
    BI_CUDA(cudaGetDevice(&dev_now));
    if (cache.k != k || cache.smem != smem || cache.dev != dev_now) {
        if (smem > 48 * 1024) BI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, n_sm = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev_now);
        cache = Cached{k, smem, dev_now, per_sm, n_sm};
    }
    const int per_sm = cache.per_sm, n_sm = cache.n_sm;
    const int64_t want = (rows + 255) / 256;
    const int64_t grid = std::min<int64_t>(want, (int64_t)n_sm * per_sm);      // persistent: one resident wave
    k<<<(unsigned)grid, 256, smem, st>>>(dX, ld, N, d, d_cut_off, d_cut_val, cut_stride, codes_pm, rows);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// point-major packed records -> feature-major code rows (on demand only; one point per thread)
__global__ void __launch_bounds__(256) unpack_codes_kernel(const uint4* __restrict__ codes_pm, int64_t Npad, int d,
                                                           uint8_t* __restrict__ codes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    const uint4 v = codes_pm[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < d; ++j) codes[(int64_t)j * Npad + i] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
}

int launch_unpack_codes(const uint4* codes_pm, int64_t Npad, int d, uint8_t* codes, cudaStream_t st) {
    if (Npad == 0 || d == 0) return BARTINT_OK;
    unpack_codes_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, st>>>(codes_pm, Npad, d, codes);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// Counting mode: per-draw sums of the single-split trees without evaluating them point by point.
//   sum_i f_tree(x_i) = mu_L * (N - G[v][c]) + mu_R * G[v][c],   G[v][c] = #{i < N : code_i[v] > c}
// G comes from 1-D histograms of the packed codes (one pass over the point set, shared by all S*m trees).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hist_codes_kernel(const uint4* __restrict__ codes_pm, int64_t N, int d,
                                                         unsigned int* __restrict__ hist /* [16][128] */) {
    __shared__ unsigned int h[16 * 128];
    for (int k = threadIdx.x; k < 16 * 128; k += blockDim.x) h[k] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = codes_pm[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < d; ++j) atomicAdd(&h[j * 128 + ((w[j >> 2] >> (8 * (j & 3))) & 127u)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 16 * 128; k += blockDim.x)
        if (h[k] != 0u) atomicAdd(&hist[k], h[k]);
}

// G[v][c] = sum_{code > c} hist[v][code]  (one thread per variable; 128 codes)
__global__ void suffix_counts_kernel(const unsigned int* __restrict__ hist, int d, long long* __restrict__ G) {
    const int v = threadIdx.x;
    if (v >= d) return;
    long long run = 0;
    for (int c = 127; c >= 0; --c) {
        G[v * 128 + c] = run;                 // codes strictly above c
        run += hist[v * 128 + c];
    }
}

// 2-D histograms for the two-split trees: pair q = (a < b) of variables, hist2[q][code_a][code_b].  One CTA per
// (pair, slice of the points) with a private 128 x 128 shared-memory histogram, flushed with global atomics.
__device__ __forceinline__ void pair_of_index(int q, int d, int& a, int& b) {
    a = 0;
    while (q >= d - 1 - a) { q -= d - 1 - a; ++a; }
    b = a + 1 + q;
}
__host__ __device__ __forceinline__ int index_of_pair(int a, int b, int d) { return a * d - a * (a + 1) / 2 + (b - a - 1); }

__global__ void __launch_bounds__(256) hist2_codes_kernel(const uint4* __restrict__ codes_pm, int64_t N, int d,
                                                          unsigned int* __restrict__ hist2 /* [pairs][128][128] */) {
    extern __shared__ unsigned int h2[];             // 128 x 128
    for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x) h2[k] = 0u;
    __syncthreads();
    int a, b;
    pair_of_index((int)blockIdx.x, d, a, b);
    const int64_t per = (N + gridDim.y - 1) / gridDim.y;
    const int64_t i0 = (int64_t)blockIdx.y * per, i1 = min(N, i0 + per);
    const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(codes_pm);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const uint32_t ca = (words[4 * i + (a >> 2)] >> (8 * (a & 3))) & 127u;
        const uint32_t cb = (words[4 * i + (b >> 2)] >> (8 * (b & 3))) & 127u;
        atomicAdd(&h2[ca * 128 + cb], 1u);
    }
    __syncthreads();
    unsigned int* __restrict__ out = hist2 + (size_t)blockIdx.x * 128 * 128;
    for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x)
        if (h2[k] != 0u) atomicAdd(&out[k], h2[k]);
}

// in place: hist2[q][x][y]  ->  J[q][x][y] = #{code_a > x and code_b > y}   (one CTA of 128 threads per pair)
__global__ void __launch_bounds__(128) suffix2_counts_kernel(unsigned int* __restrict__ hist2) {
    extern __shared__ unsigned int h2[];
    unsigned int* __restrict__ g = hist2 + (size_t)blockIdx.x * 128 * 128;
    for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x) h2[k] = g[k];
    __syncthreads();
    {   // along y, row x = threadIdx.x: E[x][y] = sum_{y' > y} h[x][y']
        unsigned int* row = h2 + threadIdx.x * 128;
        unsigned int run = 0;
        for (int y = 127; y >= 0; --y) { const unsigned int t = row[y]; row[y] = run; run += t; }
    }
    __syncthreads();
    {   // along x, column y = threadIdx.x: J[x][y] = sum_{x' > x} E[x'][y]
        unsigned int run = 0;
        for (int x = 127; x >= 0; --x) { const unsigned int t = h2[x * 128 + threadIdx.x]; h2[x * 128 + threadIdx.x] = run; run += t; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 128 * 128; k += blockDim.x) g[k] = h2[k];
}

// one warp per draw: lanes stride over the draw's trees, fixed-order shuffle reduction, result ADDED to draw_sum.
// Single-split trees (3 nodes) from the 1-D counts G; with level2 also two-split trees (5 nodes):
//   child on the LEFT  of the root: leaves (LL, LR, R):  n_R = G1,  n_LR = G2 - J,  n_LL = N - G1 - n_LR
//   child on the RIGHT of the root: leaves (L, RL, RR):  n_L = N - G1,  n_RL = G1 - J,  n_RR = J
// with G1 = #{code[v1] > c1}, G2 = #{code[v2] > c2}, J = #{both}  (same variable: J = G[v][max(c1, c2)]).
__global__ void __launch_bounds__(256) count_draw_sums_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                                              const int32_t* __restrict__ tree_leaf_off,
                                                              const double* __restrict__ leaf_mu,
                                                              const long long* __restrict__ G,
                                                              const unsigned int* __restrict__ J2, int level2, int d,
                                                              long long N, int m, int draw_begin, int draw_end,
                                                              double* __restrict__ draw_sum) {
    const int s = draw_begin + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= draw_end) return;
    double acc = 0.0;
    for (int t = lane; t < m; t += 32) {
        const int64_t k = (int64_t)s * m + t;
        const int32_t base = tree_off[k];
        const int32_t cnt = tree_off[k + 1] - base;
        const int32_t lb = tree_leaf_off[k];
        if (cnt == 3) {                                      // exactly one split: root + two leaves
            const uint2 r = nodes[base];
            const long long g = G[(r.x & 0xFFFFu) * 128 + (r.x >> 16)];
            acc += leaf_mu[lb] * (double)(N - g) + leaf_mu[lb + 1] * (double)g;
        } else if (cnt == 5 && level2) {                     // two splits (J2 exists whenever d >= 2)
            const uint2 r0 = nodes[base];
            const uint2 n1 = nodes[base + 1];
            const bool child_left = (n1.x & 0xFFFFu) != LEAF_VAR;
            const uint2 ch = child_left ? n1 : nodes[base + r0.y];
            const int v1 = (int)(r0.x & 0xFFFFu), c1 = (int)(r0.x >> 16), v2 = (int)(ch.x & 0xFFFFu), c2 = (int)(ch.x >> 16);
            // the planner left this tree in the evaluated part if its two code words fit no DUAL table (d > 12 only)
            if (!tg_dual_able((1u << (v1 >> 2)) | (1u << (v2 >> 2)), tg_regs_for(d))) continue;
            const long long g1 = G[v1 * 128 + c1], g2 = G[v2 * 128 + c2];
            long long j;
            if (v1 == v2) j = G[v1 * 128 + max(c1, c2)];
            else if (v1 < v2) j = J2[((size_t)index_of_pair(v1, v2, d) * 128 + c1) * 128 + c2];
            else j = J2[((size_t)index_of_pair(v2, v1, d) * 128 + c2) * 128 + c1];
            if (child_left) {
                const long long n_lr = g2 - j;
                acc += leaf_mu[lb] * (double)(N - g1 - n_lr) + leaf_mu[lb + 1] * (double)n_lr + leaf_mu[lb + 2] * (double)g1;
            } else {
                acc += leaf_mu[lb] * (double)(N - g1) + leaf_mu[lb + 1] * (double)(g1 - j) + leaf_mu[lb + 2] * (double)j;
            }
        }
    }
    acc = warp_sum_misc(acc);
    if (lane == 0) draw_sum[s - draw_begin] += acc;
}

int launch_count_single_splits(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                               double* d_draw_sum, cudaStream_t st) {
    const int ndraws = draw_end - draw_begin;
    if (ndraws <= 0) return BARTINT_OK;
    const int d = p->d;
    const int pairs = (f->count_level >= 2 && d >= 2) ? d * (d - 1) / 2 : 0;
    unsigned int* d_hist = nullptr;
    const size_t b1 = 16 * 128 * (sizeof(unsigned int) + sizeof(long long));
    const size_t b2 = (size_t)pairs * 128 * 128 * sizeof(unsigned int);
    BI_CUDA(cudaMallocAsync((void**)&d_hist, b1 + b2, st));
    long long* d_G = reinterpret_cast<long long*>(d_hist + 16 * 128);
    unsigned int* d_J2 = pairs ? reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(d_hist) + b1) : nullptr;
    BI_CUDA(cudaMemsetAsync(d_hist, 0, 16 * 128 * sizeof(unsigned int), st));
    if (pairs) BI_CUDA(cudaMemsetAsync(d_J2, 0, b2, st));
    if (p->N > 0) {
        const int64_t want = (p->N + 255) / 256;
        hist_codes_kernel<<<(unsigned)std::min<int64_t>(want, 148 * 4), 256, 0, st>>>(p->codes_pm, p->N, d, d_hist);
        if (pairs) {
            // per launch, not once per process: the attribute belongs to the current device's copy of the kernel
            BI_CUDA(cudaFuncSetAttribute(hist2_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 4));
            BI_CUDA(cudaFuncSetAttribute(suffix2_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 4));
            const int slices = (int)std::max<int64_t>(1, std::min<int64_t>((148 * 3 + pairs - 1) / pairs, (p->N + 4095) / 4096));
            hist2_codes_kernel<<<dim3((unsigned)pairs, (unsigned)slices), 256, 128 * 128 * 4, st>>>(p->codes_pm, p->N, d, d_J2);
        }
    }
    suffix_counts_kernel<<<1, 32, 0, st>>>(d_hist, d, d_G);
    if (pairs) suffix2_counts_kernel<<<(unsigned)pairs, 128, 128 * 128 * 4, st>>>(d_J2);
    count_draw_sums_kernel<<<(unsigned)((ndraws * 32 + 255) / 256), 256, 0, st>>>(
        f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off, f->dev.leaf_mu, d_G, d_J2, f->count_level >= 2 ? 1 : 0, d,
        (long long)p->N, f->m, draw_begin, draw_end, d_draw_sum);
    count_launch(pairs ? 5 : 3);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(d_hist, st);
    BI_CUDA(e);
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// K0g  generate_points: Monte Carlo points drawn ON THE DEVICE from the integration measure and binned on the fly, so
// the N x d fp64 matrix never exists in HBM.  Counter-based Philox4x32-10: element (point i,
// variable j) uses counter (i_lo, i_hi, j, 0) and key (seed_lo, seed_hi) -> the stream does not depend on the launch
// shape or on how points are sharded across GPUs (a shard passes the global index of its first point).
//   uniform      x = u                                   (runif, BARTInt.R:301)
//   gaussian     x = 0.5 + Phi^-1(Phi(-.5) + u (Phi(.5) - Phi(-.5)))   N(0.5,1) truncated to [0,1] (rtnorm, :304)
//   exponential  x = -log1p(-u)                          (rexp, :307)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1;
}

__device__ __forceinline__ double draw_from_measure(int measure, uint32_t w0, uint32_t w1) {
    const double u = ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);   // 53-bit [0,1)
    if (measure == BARTINT_MEASURE_UNIFORM) return u;
    if (measure == BARTINT_MEASURE_EXPONENTIAL) return -log1p(-u);
    const double plo = 0.30853753872598690, phi = 0.69146246127401310;     // Phi(-0.5), Phi(0.5)
    return 0.5 + normcdfinv(plo + (phi - plo) * u);
}

template <bool INC>
__global__ void __launch_bounds__(256) generate_points_kernel(int measure, int64_t N, int64_t first_index, uint32_t k0,
                                                              uint32_t k1, const int32_t* __restrict__ cut_off,
                                                              const double* __restrict__ cut_val,
                                                              uint32_t* __restrict__ codes4, int64_t quads_per_row,
                                                              double* __restrict__ x_out) {
    __shared__ double s_cut_pad[258];
    __shared__ float2 s_pair[257];
    double* s_cut = s_cut_pad + 1;
    const int j = blockIdx.y;
    const int c0i = cut_off[j], nc = cut_off[j + 1] - c0i;
    load_cut_tables(cut_val, c0i, nc, s_cut, s_pair);
    __syncthreads();
    const double c0 = nc > 0 ? s_cut[0] : 0.0;
    const double span = nc > 1 ? s_cut[nc - 1] - s_cut[0] : 0.0;
    const double inv_step = span > 0.0 ? (double)(nc - 1) / span : 0.0;
    const float c0f = (float)c0, invf = (float)inv_step;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < quads_per_row;
         w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = w * 4 + k;
            if (i < N) {
                const uint64_t gi = (uint64_t)(first_index + i);
                uint32_t r0, r1;
                philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)j, 0u, k0, k1, r0, r1);
                const double x = draw_from_measure(measure, r0, r1);
                if (x_out != nullptr) x_out[(int64_t)j * N + i] = x;
                word |= bin_screened<INC>(x, s_pair, s_cut, nc, c0, inv_step, c0f, invf) << (8 * k);
            }
        }
        codes4[(int64_t)j * quads_per_row + w] = word;
    }
}

int launch_generate_points(int measure, int64_t N, int64_t first_index, uint64_t seed, int32_t d,
                           const int32_t* d_cut_off, const double* d_cut_val, int route, uint8_t* codes, int64_t Npad,
                           double* x_out, cudaStream_t st) {
    if (d == 0 || Npad == 0) return BARTINT_OK;
    const int64_t quads = Npad / 4;
    int64_t bx = (quads + 255) / 256;
    const int64_t cap = (148 * 8 * 4 + d - 1) / d;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)d);
    auto kern = route == BARTINT_ROUTE_BART ? generate_points_kernel<true> : generate_points_kernel<false>;
    kern<<<grid, 256, 0, st>>>(measure, N, first_index, (uint32_t)seed, (uint32_t)(seed >> 32), d_cut_off, d_cut_val,
                               reinterpret_cast<uint32_t*>(codes), quads, x_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// Table builder: one CTA per group, one thread per table entry.  Thread idx walks every tree of the
// group taking "right" at an internal node iff its predicate bit is set in idx, and sums the leaf
// values in tree order.  Threads < NB also copy the node descriptors; thread 0 writes the header.
// ---------------------------------------------------------------------------------------------
__global__ void build_tables_kernel(const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off,
                                    const int32_t* __restrict__ tree_leaf_off, const double* __restrict__ leaf_mu,
                                    const uint8_t* __restrict__ node_bit, const int32_t* __restrict__ tree_perm,
                                    const int32_t* __restrict__ group_tree_off, const uint32_t* __restrict__ desc,
                                    int desc_words, const uint32_t* __restrict__ hdr, const uint32_t* __restrict__ split,
                                    int gather, int nb, int rec_bytes, int tbl_off, uint8_t* __restrict__ rec_base) {
    const int64_t g = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    uint8_t* rec = rec_base + g * (int64_t)rec_bytes;
    const int32_t tb = group_tree_off[g], te = group_tree_off[g + 1];
    const uint32_t flags = hdr[g];
    const uint32_t sp = split[g];
    // header: {flags, first permuted tree position (WALK: the tree id itself), number of trees, trees in table 1}
    if (tid == 0)
        *reinterpret_cast<uint4*>(rec) = make_uint4(flags, (flags & TK_FLAG_WALK) ? (uint32_t)tree_perm[tb] : (uint32_t)tb,
                                                    (uint32_t)(te - tb), sp);
    if (flags & TK_FLAG_WALK) return;       // header only: the evaluation kernel walks this tree from the global arrays
    if ((int)tid < desc_words) reinterpret_cast<uint32_t*>(rec + TK_HDR)[tid] = desc[g * desc_words + tid];
    // which table entry / which trees this thread sums
    int32_t k0 = tb, k1 = te;
    uint32_t idx = tid;
    if (gather && !(flags & TK_FLAG_JOINT)) {           // DUAL record: two 16-entry tables
        if (tid >= 32) return;
        idx = tid & 15u;
        if (tid < 16) k1 = tb + (int32_t)sp; else k0 = tb + (int32_t)sp;
    } else if (tid >= (1u << nb)) {
        return;
    }
    double sum = 0.0;
    for (int32_t kp = k0; kp < k1; ++kp) {
        const int32_t k = tree_perm[kp];
        int32_t node = tree_off[k];
        uint2 r = nodes[node];
        while ((r.x & 0xFFFFu) != LEAF_VAR) {
            const uint32_t bit = node_bit[node];
            node += ((idx >> bit) & 1u) ? (int32_t)r.y : 1;
            r = nodes[node];
        }
        sum += leaf_mu[tree_leaf_off[k] + (int32_t)r.y];
    }
    reinterpret_cast<double*>(rec + tbl_off)[tid] = sum;      // DUAL: entries 0..15 = table 1, 16..31 = table 2
}

int launch_build_tables(const bartint_forest* f, const uint32_t* d_desc, int desc_words, const uint32_t* d_hdr,
                        const uint32_t* d_split, cudaStream_t st) {
    if (f->n_groups == 0) return BARTINT_OK;
    const unsigned threads = std::max(32u, 1u << f->nb);
    build_tables_kernel<<<(unsigned)f->n_groups, threads, 0, st>>>(
        f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off, f->dev.leaf_mu, f->dev.node_bit, f->dev.tree_perm,
        f->dev.group_tree_off, d_desc, desc_words, d_hdr, d_split, f->gather ? 1 : 0, f->nb, fast_rec_bytes(f), fast_tbl_off(f),
        f->dev.group_rec);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// K3  deterministic second-stage reductions (fixed summation order, no floating-point atomics)
// ---------------------------------------------------------------------------------------------
// CTA = 32 draws x 8 tile slices: slice y adds tiles y, y+8, ... (independent loads in flight), then the 8 slice
// sums are added in slice order.  The summation tree depends only on (n_tiles), never on the launch shape.
__global__ void __launch_bounds__(256) reduce_draw_parts_kernel(const double* __restrict__ part, int n_tiles, int ndraws,
                                                                double* __restrict__ out) {
    __shared__ double s_part[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int s = blockIdx.x * 32 + x;
    double acc = 0.0;
    if (s < ndraws)
        for (int t = y; t < n_tiles; t += 8) acc += part[(int64_t)t * ndraws + s];
    s_part[y][x] = acc;
    __syncthreads();
    if (y == 0 && s < ndraws) {
        double tot = s_part[0][x];
#pragma unroll
        for (int k = 1; k < 8; ++k) tot += s_part[k][x];
        out[s] = tot;
    }
}

__global__ void __launch_bounds__(256) copy_bytes_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int64_t n16,
                                                         unsigned char* dst_tail, const unsigned char* src_tail, int tail) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

int launch_copy_bytes(void* dst, const void* src, int64_t bytes, cudaStream_t st) {
    if (bytes <= 0) return BARTINT_OK;
    BI_REQUIRE(((uintptr_t)dst & 15) == 0 && ((uintptr_t)src & 15) == 0, BARTINT_ERR_ARG, "internal: unaligned device copy");
    const int64_t n16 = bytes / 16;
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>((n16 + 255) / 256, 1), 148 * 8);
    copy_bytes_kernel<<<blocks, 256, 0, st>>>((uint4*)dst, (const uint4*)src, n16, (unsigned char*)dst + n16 * 16,
                                              (const unsigned char*)src + n16 * 16, (int)(bytes - n16 * 16));
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

int launch_reduce_draw_parts(const double* draw_part, int n_tiles, int ndraws, double* draw_sum, cudaStream_t st) {
    if (ndraws == 0) return BARTINT_OK;
    reduce_draw_parts_kernel<<<(ndraws + 31) / 32, 256, 0, st>>>(draw_part, n_tiles, ndraws, draw_sum);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// Chan et al. pairwise merge of (n, mean, M2), applied in chunk order
__device__ __forceinline__ void chan_merge(double& n, double& mean, double& m2, double nb, double meanb, double m2b) {
    if (nb == 0.0) return;
    if (n == 0.0) { n = nb; mean = meanb; m2 = m2b; return; }
    const double tot = n + nb;
    const double delta = meanb - mean;
    mean = mean + delta * (nb / tot);
    m2 = m2 + m2b + delta * delta * (n * nb / tot);
    n = tot;
}

// Chan merge of the draw chunks in chunk order.  The merge weights nb/tot and n*nb/tot depend only on the chunk sizes,
// so they are computed once per CTA (shared memory); the per-point dependent chain is then 4 fp64 ops per chunk with
// no division, and the loads of later chunks are in flight while earlier ones are merged.
__global__ void __launch_bounds__(256) merge_point_parts_kernel(const double* __restrict__ pt_mean,
                                                                const double* __restrict__ pt_m2, int n_chunks,
                                                                int chunk_draws, int ndraws, int64_t N,
                                                                double* __restrict__ mean_out, double* __restrict__ m2_out) {
    extern __shared__ double s_w[];          // [2][n_chunks]: w1 = nb/tot, w2 = n*nb/tot  (n = draws merged so far)
    for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        const double n = (double)min(c * chunk_draws, ndraws);
        const double nb = (double)min(chunk_draws, ndraws - c * chunk_draws);
        const double tot = n + nb;
        s_w[c] = nb / tot;
        s_w[n_chunks + c] = n * nb / tot;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double mean = pt_mean[i], m2 = pt_m2[i];             // chunk 0
#pragma unroll 4
    for (int c = 1; c < n_chunks; ++c) {
        const double mb = pt_mean[(int64_t)c * N + i], m2b = pt_m2[(int64_t)c * N + i];
        const double delta = mb - mean;
        mean = fma(delta, s_w[c], mean);
        m2 = fma(delta * delta, s_w[n_chunks + c], m2 + m2b);
    }
    mean_out[i] = mean;
    m2_out[i] = m2;
}

int launch_merge_point_parts(const double* pt_mean, const double* pt_m2, int n_chunks, int chunk_draws, int ndraws,
                             int64_t N, double* mean_out, double* m2_out, cudaStream_t st) {
    if (N == 0) return BARTINT_OK;
    merge_point_parts_kernel<<<(unsigned)((N + 255) / 256), 256, 2 * sizeof(double) * (size_t)n_chunks, st>>>(
        pt_mean, pt_m2, n_chunks, chunk_draws, ndraws, N, mean_out, m2_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// merge G draw shards (other GPUs' partials, shard-major) and apply colVars * weights / colMeans scaling
__global__ void merge_shards_kernel(const double* __restrict__ mean, const double* __restrict__ m2, int G,
                                    const int32_t* __restrict__ counts, int64_t N, double scale, double y0, int scaled,
                                    const double* __restrict__ w, int64_t wlen, double* __restrict__ mean_out,
                                    double* __restrict__ var_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double n = 0.0, mu = 0.0, M2 = 0.0;
    for (int g = 0; g < G; ++g) chan_merge(n, mu, M2, (double)counts[g], mean[(int64_t)g * N + i], m2[(int64_t)g * N + i]);
    if (mean_out) mean_out[i] = scaled ? mu : scale * (mu + 0.5) + y0;
    if (var_out) {
        double v = M2 / (n - 1.0);                       // n == 1 -> 0/0 = NaN like colVars on one row
        if (!scaled) v = v * scale * scale;
        if (w) v *= (wlen == 1 ? w[0] : w[i]);
        var_out[i] = v;
    }
}

int launch_merge_shards(const double* mean, const double* m2, int G, const int32_t* d_counts, int64_t N, double scale,
                        double y0, int scaled, const double* w, int64_t wlen, double* mean_out, double* var_out,
                        cudaStream_t st) {
    if (N == 0) return BARTINT_OK;
    merge_shards_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mean, m2, G, d_counts, N, scale, y0, scaled, w,
                                                                     wlen, mean_out, var_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// rowMeans(predict(model, X)):  (ymax - ymin) * (sum/N + 0.5) + ymin
__global__ void finalize_draws_kernel(const double* __restrict__ draw_sum, int ndraws, double inv_n_is_div,
                                      double scale, double y0, int scaled, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ndraws) return;
    const double mean = draw_sum[s] / inv_n_is_div;     // N == 0 -> 0/0 = NaN like rowMeans of a 0-column matrix
    out[s] = scaled ? mean : scale * (mean + 0.5) + y0;
}

int launch_finalize_draws(const double* draw_sum, int ndraws, int64_t N, double scale, double y0, int scaled,
                          double* out, cudaStream_t st) {
    if (ndraws == 0) return BARTINT_OK;
    finalize_draws_kernel<<<(ndraws + 127) / 128, 128, 0, st>>>(draw_sum, ndraws, (double)N, scale, y0, scaled, out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// K2  exact leaf-probability integral, one thread per tree (S*m trees, a few leaves each).
// Restates fillProbabilityForNode (BARTInt.R:32-72) + terminalProbability (:10-30) + the leaf sum of
// singleTreeSum (:165-173).  The box [lo, hi]^d lives in a per-thread column of a global scratch
// (interleaved so that warp accesses to the same variable coalesce); the DFS stack keeps, per level,
// what is needed to undo the box narrowing and the conditional probability of the child currently
// being visited, so P(leaf) is multiplied leaf-first going up exactly like the reference.
// ---------------------------------------------------------------------------------------------
constexpr int K2_MAX_DEPTH = 64;

__device__ __forceinline__ double pexp1(double q) {   // stats::pexp(q), rate 1
    return q <= 0.0 ? 0.0 : -expm1(-q);
}

__device__ __forceinline__ void child_probs(int measure, double lo, double hi, double rule, double& pl, double& pr) {
    if (measure == BARTINT_MEASURE_UNIFORM) {                       // :46-48
        pl = (rule - lo) / (hi - lo);
        pr = (hi - rule) / (hi - lo);
    } else if (measure == BARTINT_MEASURE_GAUSSIAN_REFCOMPAT) {     // :49-52 (as written: quirk B3)
        const double norm = normcdf(hi - 0.5) - normcdf(lo - 0.5);
        pl = (normcdf(rule) - normcdf(lo)) / norm;
        pr = 1.0 - pl;
    } else if (measure == BARTINT_MEASURE_GAUSSIAN) {               // corrected: the N(0.5, 1) mass on both sides
        const double norm = normcdf(hi - 0.5) - normcdf(lo - 0.5);
        pl = (normcdf(rule - 0.5) - normcdf(lo - 0.5)) / norm;
        pr = (normcdf(hi - 0.5) - normcdf(rule - 0.5)) / norm;
    } else {                                                        // :53-56
        const double norm = pexp1(hi) - pexp1(lo);
        pl = (pexp1(rule) - pexp1(lo)) / norm;
        pr = 1.0 - pl;
    }
}

__global__ void __launch_bounds__(128) exact_tree_integral_kernel(
    const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off, const int32_t* __restrict__ tree_leaf_off,
    const double* __restrict__ leaf_mu, const int32_t* __restrict__ cut_off, const double* __restrict__ cut_val,
    int d, int64_t n_trees, int measure, const double* __restrict__ lo0, const double* __restrict__ hi0,
    double* __restrict__ box, double* __restrict__ tree_int) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double* blo = box + tid;                       // blo[j * nthreads], bhi = blo + d * nthreads
    double* bhi = box + (int64_t)d * nthreads + tid;
    for (int j = 0; j < d; ++j) {   // default box: [0,1]^d, or [0,Inf)^d for the exponential measure (BARTInt.R:151-155)
        blo[(int64_t)j * nthreads] = lo0 ? lo0[j] : 0.0;
        bhi[(int64_t)j * nthreads] = hi0 ? hi0[j] : (measure == BARTINT_MEASURE_EXPONENTIAL ? INFINITY : 1.0);
    }

    int32_t st_var[K2_MAX_DEPTH];
    int32_t st_node[K2_MAX_DEPTH];
    uint8_t st_phase[K2_MAX_DEPTH];
    double st_rule[K2_MAX_DEPTH], st_old[K2_MAX_DEPTH], st_cp[K2_MAX_DEPTH], st_pr[K2_MAX_DEPTH];

    for (int64_t k = tid; k < n_trees; k += nthreads) {
        const int32_t base = tree_off[k];
        const double* mu = leaf_mu + tree_leaf_off[k];
        double total = 0.0;
        int sp = 0;
        int32_t node = base;
        while (true) {
            const uint2 r = nodes[node];
            const uint32_t v = r.x & 0xFFFFu;
            if (v != LEAF_VAR) {
                const double rule = cut_val[cut_off[v] + (int)(r.x >> 16)];       // :45
                const double lo = blo[(int64_t)v * nthreads], hi = bhi[(int64_t)v * nthreads];
                double pl, pr;
                child_probs(measure, lo, hi, rule, pl, pr);
                st_var[sp] = (int32_t)v; st_node[sp] = node + (int32_t)r.y; st_phase[sp] = 0;
                st_rule[sp] = rule; st_old[sp] = hi; st_cp[sp] = pl; st_pr[sp] = pr;
                ++sp;
                bhi[(int64_t)v * nthreads] = rule;                               // :62  leftCut = (lo, rule)
                node = node + 1;
                continue;
            }
            // leaf: P = prob(leaf) * prob(parent) * ... excluding the root (:22-27); stump -> 1 (:67-69)
            double p = 1.0;
            if (sp > 0) {
                p = st_cp[sp - 1];
                for (int a = sp - 2; a >= 0; --a) p = p * st_cp[a];
            }
            total = total + p * mu[r.y];                                          // :172
            // backtrack to the innermost ancestor whose right subtree is still to visit
            bool found = false;
            while (sp > 0) {
                const int a = sp - 1;
                const int32_t v2 = st_var[a];
                if (st_phase[a] == 0) {
                    bhi[(int64_t)v2 * nthreads] = st_old[a];                      // restore hi
                    st_old[a] = blo[(int64_t)v2 * nthreads];
                    blo[(int64_t)v2 * nthreads] = st_rule[a];                     // :64  rightCut = (rule, hi)
                    st_phase[a] = 1;
                    st_cp[a] = st_pr[a];
                    node = st_node[a];
                    found = true;
                    break;
                }
                blo[(int64_t)v2 * nthreads] = st_old[a];                          // restore lo
                --sp;
            }
            if (!found) break;
        }
        tree_int[k] = total;
    }
}

// posteriorSum (BARTInt.R:198): sum of the m single-tree sums of one draw, in tree order
__global__ void draw_integral_kernel(const double* __restrict__ tree_int, int m, int n_draws, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_draws) return;
    double acc = 0.0;
    for (int t = 0; t < m; ++t) acc += tree_int[(int64_t)s * m + t];
    out[s] = acc;
}

int launch_exact_integrals(const bartint_forest* f, int measure, const double* d_lo, const double* d_hi, int n_draws,
                           double* d_out, cudaStream_t st) {
    if (n_draws == 0) return BARTINT_OK;
    BI_REQUIRE(f->max_depth <= K2_MAX_DEPTH, BARTINT_ERR_UNSUPPORTED,
               "exact integral supports trees up to depth %d (forest has depth %d)", K2_MAX_DEPTH, f->max_depth);
    const int64_t n_trees = (int64_t)n_draws * f->m;
    int blocks = (int)std::min<int64_t>((n_trees + 127) / 128, 148 * 4);
    const int64_t nthreads = (int64_t)blocks * 128;
    double *box = nullptr, *tree_int = nullptr;
    BI_CUDA(cudaMallocAsync((void**)&box, sizeof(double) * (size_t)(2 * std::max(f->d, 1)) * (size_t)nthreads, st));
    BI_CUDA(cudaMallocAsync((void**)&tree_int, sizeof(double) * (size_t)n_trees, st));
    exact_tree_integral_kernel<<<blocks, 128, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off,
                                                       f->dev.leaf_mu, f->dev.cut_off, f->dev.cut_val, f->d, n_trees,
                                                       measure, d_lo, d_hi, box, tree_int);
    count_launch();
    BI_CUDA(cudaGetLastError());
    draw_integral_kernel<<<(n_draws + 127) / 128, 128, 0, st>>>(tree_int, f->m, n_draws, d_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    BI_CUDA(cudaFreeAsync(box, st));
    BI_CUDA(cudaFreeAsync(tree_int, st));
    return BARTINT_OK;
}

// BARTInt.R:260-262: rescale, mean, sd (n-1).  One CTA; fixed-order tree reduction in shared memory.
__global__ void __launch_bounds__(256) finalize_integrals_kernel(const double* __restrict__ in, int n, double y_min,
                                                                 double y_max, double* __restrict__ rescaled,
                                                                 double* __restrict__ mean_sd) {
    __shared__ double s_red[256];
    __shared__ double s_mean;
    const int t = threadIdx.x;
    double acc = 0.0;
    for (int i = t; i < n; i += 256) {
        const double v = (in[i] + 0.5) * (y_max - y_min) + y_min;     // :260
        if (rescaled) rescaled[i] = v;
        acc += v;
    }
    s_red[t] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) s_red[t] += s_red[t + w];
        __syncthreads();
    }
    if (t == 0) s_mean = s_red[0] / (double)n;                        // :261
    __syncthreads();
    const double mean = s_mean;
    acc = 0.0;
    for (int i = t; i < n; i += 256) {
        const double v = (in[i] + 0.5) * (y_max - y_min) + y_min;
        acc += (v - mean) * (v - mean);
    }
    __syncthreads();
    s_red[t] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) s_red[t] += s_red[t + w];
        __syncthreads();
    }
    if (t == 0) {
        mean_sd[0] = mean;
        mean_sd[1] = sqrt(s_red[0] / (double)(n - 1));                 // :262  (n == 1 -> NaN like R)
    }
}

// bartMean.R:80,82: mean(pred) (= mean of the row means, all rows have equal length) and the variance of the row means
__global__ void __launch_bounds__(256) draw_summary_kernel(const double* __restrict__ rm, int n, double* __restrict__ out) {
    __shared__ double s_red[256];
    __shared__ double s_mean;
    const int t = threadIdx.x;
    double acc = 0.0;
    for (int i = t; i < n; i += 256) acc += rm[i];
    s_red[t] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) s_red[t] += s_red[t + w];
        __syncthreads();
    }
    if (t == 0) s_mean = s_red[0] / (double)n;
    __syncthreads();
    const double mean = s_mean;
    acc = 0.0;
    for (int i = t; i < n; i += 256) acc += (rm[i] - mean) * (rm[i] - mean);
    __syncthreads();
    s_red[t] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) s_red[t] += s_red[t + w];
        __syncthreads();
    }
    if (t == 0) { out[0] = mean; out[1] = s_red[0] / (double)(n - 1); }
}

int launch_draw_summary(const double* d_draw_mean, int ndraws, double* d_mean_var, cudaStream_t st) {
    draw_summary_kernel<<<1, 256, 0, st>>>(d_draw_mean, ndraws, d_mean_var);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// which(var == max(var)) (BARTInt.R:315,387; bartMean.R:51): index of the first maximum and the number of ties.  One CTA;
// any NaN makes the result {-1, 0} (R: max() is NaN and the tie set is empty).
__global__ void __launch_bounds__(256) argmax_ties_kernel(const double* __restrict__ v, int64_t n, int64_t* __restrict__ out) {
    __shared__ double s_max[256];
    __shared__ long long s_idx[256];
    __shared__ long long s_cnt[256];
    __shared__ int s_nan;
    const int t = threadIdx.x;
    if (t == 0) s_nan = 0;
    __syncthreads();
    double best = -INFINITY;
    long long bi = -1;
    for (int64_t i = t; i < n; i += 256) {
        const double x = v[i];
        if (x != x) s_nan = 1;
        if (x > best || bi < 0) { best = x; bi = i; }
    }
    s_max[t] = best; s_idx[t] = bi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w && s_idx[t + w] >= 0 &&
            (s_idx[t] < 0 || s_max[t + w] > s_max[t] || (s_max[t + w] == s_max[t] && s_idx[t + w] < s_idx[t]))) {
            s_max[t] = s_max[t + w]; s_idx[t] = s_idx[t + w];
        }
        __syncthreads();
    }
    const double m = s_max[0];
    long long cnt = 0;
    for (int64_t i = t; i < n; i += 256) cnt += (v[i] == m);
    s_cnt[t] = cnt;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) s_cnt[t] += s_cnt[t + w];
        __syncthreads();
    }
    if (t == 0) {
        out[0] = (s_nan || n == 0) ? -1 : s_idx[0];
        out[1] = (s_nan || n == 0) ? 0 : s_cnt[0];
    }
}

int launch_argmax_ties(const double* d_v, int64_t n, int64_t* d_out, cudaStream_t st) {
    argmax_ties_kernel<<<1, 256, 0, st>>>(d_v, n, d_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// Batch design: the k candidates of largest posterior variance, i.e. order(var, decreasing =
// TRUE)[1:k] with ties in index order and NaN last -- the selection src/BARTInt.R:315,387 / bartMean.R:51 make for k = 1.
// One CTA; every thread keeps the best not-yet-taken element of its strided share in registers, a round is one
// block-wide arg-max (value, then lower index), and only the winner's owner rescans its share.
__global__ void __launch_bounds__(1024) topk_kernel(const double* __restrict__ v, int64_t n, int k, int64_t* __restrict__ idx_out,
                                                    double* __restrict__ val_out) {
    __shared__ double s_val[32];
    __shared__ long long s_idx[32];
    __shared__ long long s_win;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double taken_v = INFINITY;        // the last element of this thread's share that was output (share order: value desc, index asc)
    long long taken_i = -1;
    double best = -INFINITY;
    long long bi = -1;
    auto rescan = [&]() {             // best element of the share that comes AFTER (taken_v, taken_i) in the output order
        best = -INFINITY; bi = -1;
        for (int64_t i = t; i < n; i += 1024) {
            const double x = v[i];
            if (x != x) continue;                                                   // NaN: never selected
            const bool after = x < taken_v || (x == taken_v && (long long)i > taken_i);
            if (after && (bi < 0 || x > best)) { best = x; bi = i; }                // first maximum = lowest index
        }
    };
    rescan();
    for (int r = 0; r < k; ++r) {
        double bv = best;
        long long bx = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long ox = __shfl_xor_sync(0xffffffffu, bx, o);
            if (ox >= 0 && (bx < 0 || ov > bv || (ov == bv && ox < bx))) { bv = ov; bx = ox; }
        }
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bx; }
        __syncthreads();
        if (warp == 0) {
            bv = s_val[lane]; bx = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const long long ox = __shfl_xor_sync(0xffffffffu, bx, o);
                if (ox >= 0 && (bx < 0 || ov > bv || (ov == bv && ox < bx))) { bv = ov; bx = ox; }
            }
            if (lane == 0) {
                s_win = bx;
                idx_out[r] = bx;
                if (val_out) val_out[r] = bx >= 0 ? bv : NAN;
            }
        }
        __syncthreads();
        const long long win = s_win;
        if (win >= 0 && win == bi) { taken_v = best; taken_i = bi; rescan(); }
        __syncthreads();
    }
}

int launch_topk(const double* d_v, int64_t n, int k, int64_t* d_idx, double* d_val, cudaStream_t st) {
    if (k <= 0) return BARTINT_OK;
    topk_kernel<<<1, 1024, 0, st>>>(d_v, n, k, d_idx, d_val);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

int launch_finalize_integrals(const double* d_in, int n, double y_min, double y_max, double* d_rescaled,
                              double* d_mean_sd, cudaStream_t st) {
    finalize_integrals_kernel<<<1, 256, 0, st>>>(d_in, n, y_min, y_max, d_rescaled, d_mean_sd);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
