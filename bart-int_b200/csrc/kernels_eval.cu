// K1  posterior ensemble evaluation with fused reductions (the hot kernel of the path).
//
// Replaces  predict(model, X)  +  colVars  +  rowMeans  (BARTInt.R:312-314, bartMean.R:47-50,74-82):
// for every draw s, tree t and point i find the leaf, add its mu (fp64), and reduce on the fly to
//   draw_part[tile][s]   = sum over the tile's points of  sum_t mu          (-> rowMeans)
//   pt_mean/pt_m2[c][i]  = mean / sum of squared deviations over draw chunk c (-> colVars)
// Two kernels:
//   eval_table_kernel  the fast path.  Forest = per-group records (node descriptors + joint leaf
//                      table) streamed through shared memory with 1-D bulk TMA copies and an
//                      mbarrier ring; points = uint8 cut codes, 4 per 32-bit word, feature-major
//                      in shared memory; node test = one byte-SIMD add for 4 points; leaf values of
//                      a whole group = one fp64 table lookup per point.
//   eval_walk_kernel   generic fallback (any tree depth / up to 255 cuts): per-lane node walk from
//                      L1/L2, also the producer of leaf ordinals for the bit-exactness check.
#include <algorithm>

#include "common.cuh"

namespace bartint {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a lost TMA completion traps (launch error) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// work decomposition
// ---------------------------------------------------------------------------------------------
constexpr int WK_THREADS = 256;   // walk kernel: one point per thread

EvalPlan make_plan(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                   int32_t draw_end, bool want_moments) {
    (void)want_moments;
    EvalPlan pl{};
    pl.draw_begin = draw_begin;
    pl.draw_end = draw_end;
    const int ndraws = draw_end - draw_begin;
    pl.use_table = f->table_ok && !(flags & BARTINT_FORCE_WALK) && p->d <= f->d;
    const int tile = pl.use_table ? TK_TILE : WK_THREADS;
    pl.n_tiles = (int)((p->N + tile - 1) / tile);
    if (pl.n_tiles < 1) pl.n_tiles = 1;
    // CTAs resident at once: 148 SMs x 2 (table kernel) or x 8 (walk kernel).  Split the draw range into chunks so
    // the grid is several full waves; the tail wave then costs little.  Chunks only add tiny second-stage merges.
    const int resident = 148 * (pl.use_table ? 2 : 8);
    int best_chunks = 1;
    double best_eff = -1.0;
    const int max_chunks = std::max(1, std::min(ndraws, pl.use_table ? 64 : 16));
    for (int c = 1; c <= max_chunks; ++c) {
        const int cd = (ndraws + c - 1) / c;
        const int real_c = (ndraws + cd - 1) / cd;
        if (real_c != c) continue;
        const double ctas = (double)pl.n_tiles * c;
        const double waves = ctas / resident;
        double eff = waves / std::ceil(waves);
        // each extra chunk re-loads the tile's codes and re-fills the pipeline: mild preference for few chunks
        eff -= 0.002 * c;
        if (cd < 4 && c > 1) continue;
        if (eff > best_eff + 1e-9) { best_eff = eff; best_chunks = c; }
    }
    pl.n_chunks = best_chunks;
    pl.chunk_draws = std::max(1, (ndraws + best_chunks - 1) / best_chunks);
    return pl;
}

// ---------------------------------------------------------------------------------------------
// generic walk kernel
// ---------------------------------------------------------------------------------------------
template <bool MOMENTS>
__global__ void __launch_bounds__(WK_THREADS) eval_walk_kernel(
    const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off, const int32_t* __restrict__ tree_leaf_off,
    const double* __restrict__ leaf_mu, const uint8_t* __restrict__ codes, int64_t Npad, int64_t N, int m,
    int draw_begin, int draw_end, int chunk_draws, double* __restrict__ draw_part, double* __restrict__ pt_mean,
    double* __restrict__ pt_m2, int32_t* __restrict__ leaf_out, double* __restrict__ full_pred, double fp_scale,
    double fp_y0, int fp_scaled) {
    __shared__ double s_warp[WK_THREADS / 32];
    const int tile = blockIdx.x, chunk = blockIdx.y;
    const int64_t i = (int64_t)tile * WK_THREADS + threadIdx.x;
    const bool valid = i < N;
    const int ndraws = draw_end - draw_begin;
    const int s0 = draw_begin + chunk * chunk_draws;
    const int s1 = min(s0 + chunk_draws, draw_end);
    const uint8_t* __restrict__ my_codes = codes + i;   // + var * Npad ; the padded tail is readable (zeros)
    double piv = 0.0, S1 = 0.0, S2 = 0.0;
    for (int s = s0; s < s1; ++s) {
        double acc = 0.0;
        for (int t = 0; t < m; ++t) {
            const int64_t k = (int64_t)s * m + t;
            int32_t node = tree_off[k];
            uint2 r = __ldg(&nodes[node]);
            while ((r.x & 0xFFFFu) != LEAF_VAR) {
                const uint32_t c = my_codes[(int64_t)(r.x & 0xFFFFu) * Npad];
                node += (c > (r.x >> 16)) ? (int32_t)r.y : 1;     // right iff code > cut index  <=>  x > cut value
                r = __ldg(&nodes[node]);
            }
            acc += __ldg(&leaf_mu[tree_leaf_off[k] + (int32_t)r.y]);
            if (leaf_out != nullptr && valid) leaf_out[(k - (int64_t)draw_begin * m) * N + i] = (int32_t)r.y;
        }
        if (full_pred != nullptr && valid)
            full_pred[(int64_t)(s - draw_begin) + (int64_t)ndraws * i] = fp_scaled ? acc : fp_scale * (acc + 0.5) + fp_y0;
        if (MOMENTS) {
            if (s == s0) piv = acc;
            const double dlt = acc - piv;
            S1 += dlt;
            S2 = fma(dlt, dlt, S2);
        }
        double v = warp_sum(valid ? acc : 0.0);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < WK_THREADS / 32; ++w) tot += s_warp[w];
            draw_part[(int64_t)tile * ndraws + (s - draw_begin)] = tot;
        }
        __syncthreads();
    }
    if (MOMENTS && valid) {
        const double n = (double)(s1 - s0);
        pt_mean[(int64_t)chunk * N + i] = piv + S1 / n;
        pt_m2[(int64_t)chunk * N + i] = S2 - S1 * S1 / n;
    }
}

int launch_eval_walk(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                     double* pt_mean, double* pt_m2, int32_t* leaf_out, double* full_pred, double fp_scale,
                     double fp_y0, int fp_scaled, cudaStream_t st) {
    dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.n_chunks);
    auto kern = pt_mean != nullptr ? eval_walk_kernel<true> : eval_walk_kernel<false>;
    kern<<<grid, WK_THREADS, 0, st>>>(f->dev.nodes, f->dev.tree_off, f->dev.tree_leaf_off, f->dev.leaf_mu, p->codes,
                                      p->Npad, p->N, f->m, plan.draw_begin, plan.draw_end, plan.chunk_draws, draw_part,
                                      pt_mean, pt_m2, leaf_out, full_pred, fp_scale, fp_y0, fp_scaled);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// table kernel
// ---------------------------------------------------------------------------------------------
constexpr int TK_GS = 8;       // groups per pipeline stage
constexpr int TK_NSTAGE = 3;   // stages in flight

// Predicate bytes of one group for one quad word stream.  After the NB steps, byte k of A holds, in its
// top NB bits, the predicates of nodes 0..NB-1 for point k of the quad (node j at bit 8-NB+j).
// SHARED by the evaluation kernel and the leaf-decoding kernel so the bit-exactness check exercises
// exactly the compare the fast path uses.
__device__ __forceinline__ uint32_t pred_step(uint32_t A, uint32_t codes4, uint32_t bias4) {
    return (A >> 1) | ((codes4 + bias4) & 0x80808080u);
}

struct TableParams {
    const uint8_t* group_rec;
    const int32_t* draw_group_off;
    const uint8_t* codes;
    int64_t Npad, N;
    int d;
    int draw_begin, draw_end, chunk_draws;
    double* draw_part;
    double* pt_mean;
    double* pt_m2;
    double* full_pred;
    double fp_scale, fp_y0;
    int fp_scaled;
};

template <int NB>
__host__ __device__ constexpr int tk_stage_bytes() { return TK_GS * tk_rec_bytes(NB); }

template <int NB>
size_t tk_smem_bytes(int d) {
    return (size_t)d * TK_TILE + (size_t)TK_NSTAGE * tk_stage_bytes<NB>() + 2 * TK_GS * (TK_THREADS / 32) * sizeof(double) +
           (TK_NSTAGE + 1) * sizeof(uint64_t);
}

template <int NB, bool MOMENTS, bool FULL>
__global__ void __launch_bounds__(TK_THREADS, 2) eval_table_kernel(const TableParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int REC = tk_rec_bytes(NB);
    constexpr int STAGE = tk_stage_bytes<NB>();
    constexpr int NWARP = TK_THREADS / 32;
    unsigned char* s_codes = smem;                                        // d rows of TK_TILE code bytes
    unsigned char* s_stage = smem + (size_t)prm.d * TK_TILE;              // TK_NSTAGE stage buffers
    double* s_wsum = reinterpret_cast<double*>(s_stage + TK_NSTAGE * STAGE);   // [2][TK_GS][NWARP]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_wsum + 2 * TK_GS * NWARP); // [TK_NSTAGE] full + [1] codes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, chunk = blockIdx.y;
    const int ndraws = prm.draw_end - prm.draw_begin;
    const int s0 = prm.draw_begin + chunk * prm.chunk_draws;
    const int s1 = min(s0 + prm.chunk_draws, prm.draw_end);
    const int g0 = prm.draw_group_off[s0], g1 = prm.draw_group_off[s1];
    const int n_groups = g1 - g0;
    const int n_stages = (n_groups + TK_GS - 1) / TK_GS;
    const unsigned char* __restrict__ rec_g = prm.group_rec + (int64_t)g0 * REC;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b <= TK_NSTAGE; ++b) mbar_init(&s_bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        // this tile's codes: one bulk copy per variable row
        mbar_expect_tx(&s_bar[TK_NSTAGE], (uint32_t)prm.d * TK_TILE);
        for (int j = 0; j < prm.d; ++j)
            bulk_g2s(s_codes + (size_t)j * TK_TILE, prm.codes + (int64_t)j * prm.Npad + (int64_t)tile * TK_TILE, TK_TILE,
                     &s_bar[TK_NSTAGE]);
        for (int k = 0; k < TK_NSTAGE && k < n_stages; ++k) {
            const int ng = min(TK_GS, n_groups - k * TK_GS);
            mbar_expect_tx(&s_bar[k], (uint32_t)(ng * REC));
            bulk_g2s(s_stage + k * STAGE, rec_g + (int64_t)k * STAGE, (uint32_t)(ng * REC), &s_bar[k]);
        }
    }

    // validity of this thread's 8 points: quad q covers tile points (q*THREADS + tid)*4 .. +3
    uint32_t vmask = 0;
    const int64_t tile_base = (int64_t)tile * TK_TILE;
#pragma unroll
    for (int q = 0; q < TK_Q; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (tile_base + (int64_t)(q * TK_THREADS + tid) * 4 + k < prm.N) vmask |= 1u << (q * 4 + k);

    double acc[TK_P];
    double piv[MOMENTS ? TK_P : 1], S1[MOMENTS ? TK_P : 1], S2[MOMENTS ? TK_P : 1];
#pragma unroll
    for (int p = 0; p < TK_P; ++p) acc[p] = 0.0;
#pragma unroll
    for (int p = 0; p < (MOMENTS ? TK_P : 1); ++p) { piv[p] = 0.0; S1[p] = 0.0; S2[p] = 0.0; }

    mbar_wait(&s_bar[TK_NSTAGE], 0);
    const uint32_t* __restrict__ my_codes = reinterpret_cast<const uint32_t*>(s_codes) + tid;

    int draws_done = 0;   // uniform across the CTA
    for (int k = 0; k < n_stages; ++k) {
        const int b = k % TK_NSTAGE;
        mbar_wait(&s_bar[b], (uint32_t)((k / TK_NSTAGE) & 1));
        const int ng = min(TK_GS, n_groups - k * TK_GS);
        const unsigned char* stage = s_stage + b * STAGE;
        const int par = k & 1;
        const int stage_first_draw = draws_done;
        for (int g = 0; g < ng; ++g) {
            const unsigned char* rec = stage + g * REC;
            const uint32_t flags = *reinterpret_cast<const uint32_t*>(rec);
            const uint2* __restrict__ desc = reinterpret_cast<const uint2*>(rec + TK_HDR);
            const double* __restrict__ tbl = reinterpret_cast<const double*>(rec + tk_tbl_off(NB));
            if (flags & TK_FLAG_WALK) {
                // rare: a tree too deep for NB predicate bits, walked per point from its record in shared memory
                const uint2* __restrict__ wn = reinterpret_cast<const uint2*>(rec + TK_HDR);
                const uint32_t wcnt = reinterpret_cast<const uint32_t*>(rec)[2];
                const double* __restrict__ wmu = reinterpret_cast<const double*>(rec + TK_HDR + 8 * wcnt);
#pragma unroll 1
                for (int p = 0; p < TK_P; ++p) {
                    const uint32_t bo = (uint32_t)(((p >> 2) * TK_THREADS + tid) * 4 + (p & 3));
                    uint32_t node = 0;
                    uint2 r = wn[0];
                    while ((r.x & 0xFFFFu) != LEAF_VAR) {
                        const uint32_t c = s_codes[(r.x & 0xFFFFu) * TK_TILE + bo];
                        node += (c > (r.x >> 16)) ? r.y : 1u;
                        r = wn[node];
                    }
                    const double v = wmu[r.y];
#pragma unroll
                    for (int pp = 0; pp < TK_P; ++pp)
                        if (pp == p) acc[pp] += v;      // keeps acc[] in registers (no dynamic indexing)
                }
            } else {
                uint32_t A[TK_Q];
#pragma unroll
                for (int q = 0; q < TK_Q; ++q) A[q] = 0u;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const uint2 dj = desc[j];                                    // warp-uniform: smem broadcast
                    const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(
                        reinterpret_cast<const unsigned char*>(my_codes) + dj.x);
#pragma unroll
                    for (int q = 0; q < TK_Q; ++q) A[q] = pred_step(A[q], row[q * TK_THREADS], dj.y);
                }
#pragma unroll
                for (int q = 0; q < TK_Q; ++q)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t idx = (A[q] >> (8 * kk + 8 - NB)) & ((1u << NB) - 1u);
                        acc[q * 4 + kk] += tbl[idx];
                    }
            }
            if (flags & TK_FLAG_END_DRAW) {
                // ---- draw epilogue: fold f_s(x) into the per-point moments and the per-draw sum ----
                const int s = s0 + draws_done;
                double dsum = 0.0;
#pragma unroll
                for (int p = 0; p < TK_P; ++p) {
                    const double fv = acc[p];
                    acc[p] = 0.0;
                    dsum += ((vmask >> p) & 1u) ? fv : 0.0;
                    if (MOMENTS) {
                        if (draws_done == 0) piv[p] = fv;
                        const double dlt = fv - piv[p];
                        S1[p] += dlt;
                        S2[p] = fma(dlt, dlt, S2[p]);
                    }
                    if (FULL) {
                        if ((vmask >> p) & 1u) {
                            const int64_t i = tile_base + (int64_t)((p >> 2) * TK_THREADS + tid) * 4 + (p & 3);
                            prm.full_pred[(int64_t)(s - prm.draw_begin) + (int64_t)ndraws * i] =
                                prm.fp_scaled ? fv : prm.fp_scale * (fv + 0.5) + prm.fp_y0;
                        }
                    }
                }
                dsum = warp_sum(dsum);
                if (lane == 0) s_wsum[(par * TK_GS + (draws_done - stage_first_draw)) * NWARP + warp] = dsum;
                ++draws_done;
            }
        }
        __syncthreads();   // every thread is done with stage buffer b and has published its warp sums
        if (tid == 0 && k + TK_NSTAGE < n_stages) {
            const int kn = k + TK_NSTAGE;
            const int ngn = min(TK_GS, n_groups - kn * TK_GS);
            mbar_expect_tx(&s_bar[b], (uint32_t)(ngn * REC));
            bulk_g2s(s_stage + b * STAGE, rec_g + (int64_t)kn * STAGE, (uint32_t)(ngn * REC), &s_bar[b]);
        }
        const int done_here = draws_done - stage_first_draw;
        if (tid < done_here) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) tot += s_wsum[(par * TK_GS + tid) * NWARP + w];
            prm.draw_part[(int64_t)tile * ndraws + (s0 - prm.draw_begin + stage_first_draw + tid)] = tot;
        }
    }

    if (MOMENTS) {
        const double n = (double)(s1 - s0);
#pragma unroll
        for (int p = 0; p < TK_P; ++p) {
            if ((vmask >> p) & 1u) {
                const int64_t i = tile_base + (int64_t)((p >> 2) * TK_THREADS + tid) * 4 + (p & 3);
                prm.pt_mean[(int64_t)chunk * prm.N + i] = piv[p] + S1[p] / n;
                prm.pt_m2[(int64_t)chunk * prm.N + i] = S2[p] - S1[p] * S1[p] / n;
            }
        }
    }
}

template <int NB, bool MOMENTS, bool FULL>
static int launch_table_inst(const TableParams& prm, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = tk_smem_bytes<NB>(prm.d);
    auto kern = eval_table_kernel<NB, MOMENTS, FULL>;
    static size_t configured_smem = 0;
    if (!configured || smem > configured_smem) {
        BI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
        configured_smem = 227 * 1024;
    }
    kern<<<grid, TK_THREADS, smem, st>>>(prm);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

template <int NB>
static int launch_table_nb(const TableParams& prm, dim3 grid, cudaStream_t st) {
    const bool mom = prm.pt_mean != nullptr, full = prm.full_pred != nullptr;
    if (mom && full) return launch_table_inst<NB, true, true>(prm, grid, st);
    if (mom) return launch_table_inst<NB, true, false>(prm, grid, st);
    if (full) return launch_table_inst<NB, false, true>(prm, grid, st);
    return launch_table_inst<NB, false, false>(prm, grid, st);
}

int launch_eval_table(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                      double* pt_mean, double* pt_m2, double* full_pred, double fp_scale, double fp_y0, int fp_scaled,
                      cudaStream_t st) {
    TableParams prm{};
    prm.group_rec = f->dev.group_rec;
    prm.draw_group_off = f->dev.draw_group_off;
    prm.codes = p->codes;
    prm.Npad = p->Npad;
    prm.N = p->N;
    prm.d = p->d;
    prm.draw_begin = plan.draw_begin;
    prm.draw_end = plan.draw_end;
    prm.chunk_draws = plan.chunk_draws;
    prm.draw_part = draw_part;
    prm.pt_mean = pt_mean;
    prm.pt_m2 = pt_m2;
    prm.full_pred = full_pred;
    prm.fp_scale = fp_scale;
    prm.fp_y0 = fp_y0;
    prm.fp_scaled = fp_scaled;
    dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.n_chunks);
    switch (f->nb) {
        case 4: return launch_table_nb<4>(prm, grid, st);
        case 5: return launch_table_nb<5>(prm, grid, st);
        case 6: return launch_table_nb<6>(prm, grid, st);
        case 7: return launch_table_nb<7>(prm, grid, st);
        case 8: return launch_table_nb<8>(prm, grid, st);
    }
    set_error("unsupported table bits %d", f->nb);
    return BARTINT_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// Leaf ordinals decoded from the table kernel's predicate bytes (bit-exactness check of the fast
// path's compare).  One thread per (point quad, group): recompute A with pred_step from the same
// code words and descriptors, then walk each tree of the group with the predicate bits.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_leaf_kernel(
    const uint8_t* __restrict__ group_rec, int nb, const int32_t* __restrict__ draw_group_off,
    const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off, const uint8_t* __restrict__ node_bit,
    const uint8_t* __restrict__ codes, int64_t Npad, int64_t N, int m, int draw_begin, int draw_end,
    int32_t* __restrict__ leaf_out) {
    const int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = draw_group_off[draw_begin] + blockIdx.y;
    if (g >= draw_group_off[draw_end] || quad * 4 >= N) return;
    const uint8_t* rec = group_rec + (int64_t)g * tk_rec_bytes(nb);
    const uint4 h = *reinterpret_cast<const uint4*>(rec);
    if (h.x & TK_FLAG_WALK) {
        const uint2* wn = reinterpret_cast<const uint2*>(rec + TK_HDR);
        for (int kk = 0; kk < 4; ++kk) {
            const int64_t i = quad * 4 + kk;
            if (i >= N) break;
            uint32_t node = 0;
            uint2 r = wn[0];
            while ((r.x & 0xFFFFu) != LEAF_VAR) {
                const uint32_t c = codes[(int64_t)(r.x & 0xFFFFu) * Npad + i];
                node += (c > (r.x >> 16)) ? r.y : 1u;
                r = wn[node];
            }
            leaf_out[((int64_t)h.y - (int64_t)draw_begin * m) * N + i] = (int32_t)r.y;
        }
        return;
    }
    const uint2* desc = reinterpret_cast<const uint2*>(rec + TK_HDR);
    uint32_t A = 0;
    for (int j = 0; j < nb; ++j) {
        const uint2 dj = desc[j];
        const int64_t var = dj.x / TK_TILE;
        const uint32_t c4 = *reinterpret_cast<const uint32_t*>(codes + var * Npad + quad * 4);
        A = pred_step(A, c4, dj.y);
    }
    for (int kk = 0; kk < 4; ++kk) {
        const int64_t i = quad * 4 + kk;
        if (i >= N) break;
        const uint32_t idx = (A >> (8 * kk + 8 - nb)) & ((1u << nb) - 1u);
        for (uint32_t t = 0; t < h.z; ++t) {
            const int64_t k = (int64_t)h.y + t;
            int32_t node = tree_off[k];
            uint2 r = nodes[node];
            while ((r.x & 0xFFFFu) != LEAF_VAR) {
                node += ((idx >> node_bit[node]) & 1u) ? (int32_t)r.y : 1;
                r = nodes[node];
            }
            leaf_out[(k - (int64_t)draw_begin * m) * N + i] = (int32_t)r.y;
        }
    }
}

int launch_table_leaf(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                      int32_t* leaf_out, cudaStream_t st) {
    const int n_groups = f->draw_group_off[(size_t)draw_end] - f->draw_group_off[(size_t)draw_begin];
    if (n_groups == 0 || p->N == 0) return BARTINT_OK;
    const int64_t quads = (p->N + 3) / 4;
    BI_REQUIRE(n_groups <= 65535, BARTINT_ERR_UNSUPPORTED, "leaf decoding limited to 65535 groups per call");
    dim3 grid((unsigned)((quads + 255) / 256), (unsigned)n_groups);
    table_leaf_kernel<<<grid, 256, 0, st>>>(f->dev.group_rec, f->nb, f->dev.draw_group_off, f->dev.nodes,
                                            f->dev.tree_off, f->dev.node_bit, p->codes, p->Npad, p->N, f->m,
                                            draw_begin, draw_end, leaf_out);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
