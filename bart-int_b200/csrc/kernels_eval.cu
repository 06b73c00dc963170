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
//   eval_walk_kernel   generic fallback (up to 255 cuts, any number of variables): per-lane node walk
//                      from L1/L2, also the producer of leaf ordinals for the bit-exactness check.
#include <algorithm>
#include <cstdlib>

#include "pipeline.cuh"

namespace bartint {

// ---------------------------------------------------------------------------------------------
// generic walk kernel
// ---------------------------------------------------------------------------------------------
constexpr int WK_THREADS = 256;   // one point per thread

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
__host__ __device__ constexpr int tk_groups_per_stage(int nb) { return groups_per_stage(tk_rec_bytes(nb)); }

// Predicate bits of one group, 4 points per word.  The byte-wise sum codes + (127 - cut) carries into bit 7 of a
// byte iff code > cut ("go right"); PRMT with the sign-replicating selector turns bit 7 into a full byte mask and
// one LOP3 deposits it at the node's bit position.  Bit of node j inside each byte: tk_bit0(NB) + j, chosen so that
// for NB <= 5 the byte IS the table offset idx * 8.  SHARED by the evaluation kernel and the leaf-decoding kernel
// so the bit-exactness check exercises exactly the compare the fast path uses.
__host__ __device__ constexpr int tk_bit0(int nb) { return nb <= 5 ? 3 : 8 - nb; }

__device__ __forceinline__ uint32_t pred_mask4(uint32_t codes4, uint32_t bias4) {
    // per byte: 0xFF if (code + bias) >= 128 else 0x00.  Selector nibbles 8..B = "replicate the sign bit of byte
    // 0..3" -- only reachable through PTX prmt (__byte_perm masks the selector to 3 bits).
    uint32_t m;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(codes4 + bias4), "r"(0u), "r"(0xBA98u));
    return m;
}
template <int NB>
__device__ __forceinline__ uint32_t pred_step(uint32_t A, uint32_t codes4, uint32_t bias4, int j) {
    return A | (pred_mask4(codes4, bias4) & ((1u << (tk_bit0(NB) + j)) * 0x01010101u));
}
// byte offset into the group's table for point k (0..3) of the quad
template <int NB>
__device__ __forceinline__ uint32_t table_offset(uint32_t A, int k) {
    const uint32_t byte = __byte_perm(A, 0u, 0x4440u | (uint32_t)k);
    if constexpr (NB <= 5) return byte;             // bits [3, 3+NB): idx * 8
    else return byte << (NB - 5);                   // bits [8-NB, 8): idx << (8-NB)  ->  idx * 8
}
template <int NB>
__device__ __forceinline__ uint32_t table_index(uint32_t A, int k) { return table_offset<NB>(A, k) >> 3; }

template <int NB, int Q, int T>
size_t tk_smem_bytes(int d) {
    constexpr int GS = tk_groups_per_stage(NB);
    return (size_t)d * (T * 4 * Q) + (size_t)TK_NSTAGE * GS * tk_rec_bytes(NB) + 2 * GS * (T / 32) * sizeof(double) +
           (TK_NSTAGE + 1) * sizeof(uint64_t);
}

// Q = point quads per thread, T = threads per CTA: the CTA tile is T * 4 * Q points
template <int NB, int Q, int T, bool MOMENTS, bool FULL>
__global__ void __launch_bounds__(T, (T > 256) ? 2 : ((MOMENTS || Q > 2) ? 2 : 3)) eval_table_kernel(const TableParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int P = 4 * Q;
    constexpr int TILE = T * P;
    constexpr int REC = tk_rec_bytes(NB);
    constexpr int GS = tk_groups_per_stage(NB);
    constexpr int STAGE = GS * REC;
    constexpr int NWARP = T / 32;
    unsigned char* s_codes = smem;                                        // d rows of TILE code bytes
    unsigned char* s_stage = smem + (size_t)prm.d * TILE;                 // TK_NSTAGE stage buffers
    double* s_wsum = reinterpret_cast<double*>(s_stage + TK_NSTAGE * STAGE);   // [2][GS][NWARP]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_wsum + 2 * GS * NWARP);    // [TK_NSTAGE] full + [1] codes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, chunk = blockIdx.y;
    const int ndraws = prm.draw_end - prm.draw_begin;
    const int s0 = prm.draw_begin + chunk * prm.chunk_draws;
    const int s1 = min(s0 + prm.chunk_draws, prm.draw_end);
    const int g0 = prm.draw_group_off[s0], g1 = prm.draw_group_off[s1];
    const int n_groups = g1 - g0;
    const int n_stages = (n_groups + GS - 1) / GS;
    const unsigned char* __restrict__ rec_g = prm.group_rec + (int64_t)g0 * REC;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b <= TK_NSTAGE; ++b) mbar_init(&s_bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        // this tile's codes: one bulk copy per variable row
        mbar_expect_tx(&s_bar[TK_NSTAGE], (uint32_t)prm.d * TILE);
        for (int j = 0; j < prm.d; ++j)
            bulk_g2s(s_codes + (size_t)j * TILE, prm.codes + (int64_t)j * prm.Npad + (int64_t)tile * TILE, TILE,
                     &s_bar[TK_NSTAGE]);
        for (int k = 0; k < TK_NSTAGE && k < n_stages; ++k) {
            const int ng = min(GS, n_groups - k * GS);
            mbar_expect_tx(&s_bar[k], (uint32_t)(ng * REC));
            bulk_g2s(s_stage + k * STAGE, rec_g + (int64_t)k * STAGE, (uint32_t)(ng * REC), &s_bar[k]);
        }
    }

    // validity of this thread's points: quad q covers tile points (q*THREADS + tid)*4 .. +3
    uint32_t vmask = 0;
    const int64_t tile_base = (int64_t)tile * TILE;
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (tile_base + (int64_t)(q * T + tid) * 4 + k < prm.N) vmask |= 1u << (q * 4 + k);

    double acc[P];
    double piv[MOMENTS ? P : 1], S1[MOMENTS ? P : 1], S2[MOMENTS ? P : 1];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.0;
#pragma unroll
    for (int p = 0; p < (MOMENTS ? P : 1); ++p) { piv[p] = 0.0; S1[p] = 0.0; S2[p] = 0.0; }

    mbar_wait(&s_bar[TK_NSTAGE], 0);
    const unsigned char* __restrict__ my_codes = s_codes + tid * 4;

    int draws_done = 0;   // uniform across the CTA
    for (int k = 0; k < n_stages; ++k) {
        const int b = k % TK_NSTAGE;
        mbar_wait(&s_bar[b], (uint32_t)((k / TK_NSTAGE) & 1));
        const int ng = min(GS, n_groups - k * GS);
        const unsigned char* stage = s_stage + b * STAGE;
        const int par = k & 1;
        const int stage_first_draw = draws_done;
        // the stage's group flags as two warp-uniform bit masks (one strided load + ballots per stage instead of a
        // dependent header load + branch per group)
        unsigned long long walk_mask = 0ull, end_mask = 0ull;
#pragma unroll
        for (int base = 0; base < GS; base += 32) {
            const int gl = base + lane;
            const uint32_t fl = gl < ng ? *reinterpret_cast<const uint32_t*>(stage + gl * REC) : 0u;
            walk_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & TK_FLAG_WALK) != 0u) << base;
            end_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & TK_FLAG_END_DRAW) != 0u) << base;
        }
        // plain groups (no flag) run in a tight loop; the next flagged group is found with one ffs on the masks
        auto table_group = [&](const unsigned char* rec) {
            const uint2* __restrict__ cur_desc = reinterpret_cast<const uint2*>(rec + TK_HDR);   // smem broadcast loads
            const unsigned char* __restrict__ tbl = rec + tk_tbl_off(NB);
            uint32_t A[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) A[q] = 0u;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const uint2 dj = cur_desc[j];                                // {variable, bias x4}
                const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(my_codes + dj.x * TILE);
#pragma unroll
                for (int q = 0; q < Q; ++q) A[q] = pred_step<NB>(A[q], row[q * T], dj.y, j);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    acc[q * 4 + kk] += *reinterpret_cast<const double*>(tbl + table_offset<NB>(A[q], kk));
        };
        int g = 0;
#pragma unroll 1
        while (g < ng) {
            const unsigned long long special = (walk_mask | end_mask) >> g;
            const int g_stop = special ? min(ng, g + __ffsll((long long)special) - 1) : ng;
#pragma unroll 1
            for (; g < g_stop; ++g) table_group(stage + g * REC);
            if (g >= ng) break;
            const unsigned char* rec = stage + g * REC;
            if ((walk_mask >> g) & 1ull) {
                // rare: a tree with more than NB internal nodes, walked per point from global memory (L1/L2)
                const int64_t kt = reinterpret_cast<const uint32_t*>(rec)[1];
                const uint2* __restrict__ wn = prm.nodes + prm.tree_off[kt];
                const double* __restrict__ wmu = prm.leaf_mu + prm.tree_leaf_off[kt];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    // the 4 points of a quad advance in lock-step (independent chains -> memory-level parallelism)
                    uint32_t node[4] = {0u, 0u, 0u, 0u};
                    uint2 r[4];
                    const uint32_t bo = (uint32_t)((q * T + tid) * 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) r[kk] = wn[0];
                    bool any = (r[0].x & 0xFFFFu) != LEAF_VAR;
                    while (any) {
                        any = false;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if ((r[kk].x & 0xFFFFu) != LEAF_VAR) {
                                const uint32_t c = s_codes[(r[kk].x & 0xFFFFu) * TILE + bo + kk];
                                node[kk] += (c > (r[kk].x >> 16)) ? r[kk].y : 1u;
                                r[kk] = wn[node[kk]];
                                any |= (r[kk].x & 0xFFFFu) != LEAF_VAR;
                            }
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) acc[q * 4 + kk] += wmu[r[kk].y];
                }
            } else {
                table_group(rec);
            }
            if ((end_mask >> g) & 1ull) {
                // ---- draw epilogue: fold f_s(x) into the per-point moments and the per-draw sum ----
                const int s = s0 + draws_done;
                double dsum = 0.0;
#pragma unroll
                for (int p = 0; p < P; ++p) {
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
                            const int64_t i = tile_base + (int64_t)((p >> 2) * T + tid) * 4 + (p & 3);
                            prm.full_pred[(int64_t)(s - prm.draw_begin) + (int64_t)ndraws * i] =
                                prm.fp_scaled ? fv : prm.fp_scale * (fv + 0.5) + prm.fp_y0;
                        }
                    }
                }
                dsum = warp_sum(dsum);
                if (lane == 0) s_wsum[(par * GS + (draws_done - stage_first_draw)) * NWARP + warp] = dsum;
                ++draws_done;
            }
            ++g;
        }
        __syncthreads();   // every thread is done with stage buffer b and has published its warp sums
        if (tid == 0 && k + TK_NSTAGE < n_stages) {
            const int kn = k + TK_NSTAGE;
            const int ngn = min(GS, n_groups - kn * GS);
            mbar_expect_tx(&s_bar[b], (uint32_t)(ngn * REC));
            bulk_g2s(s_stage + b * STAGE, rec_g + (int64_t)kn * STAGE, (uint32_t)(ngn * REC), &s_bar[b]);
        }
        const int done_here = draws_done - stage_first_draw;
        if (tid < done_here) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) tot += s_wsum[(par * GS + tid) * NWARP + w];
            prm.draw_part[(int64_t)tile * ndraws + (s0 - prm.draw_begin + stage_first_draw + tid)] = tot;
        }
    }

    if (MOMENTS) {
        const double n = (double)(s1 - s0);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if ((vmask >> p) & 1u) {
                const int64_t i = tile_base + (int64_t)((p >> 2) * T + tid) * 4 + (p & 3);
                prm.pt_mean[(int64_t)chunk * prm.N + i] = piv[p] + S1[p] / n;
                prm.pt_m2[(int64_t)chunk * prm.N + i] = S2[p] - S1[p] * S1[p] / n;
            }
        }
    }
}

// ---- instantiation table -----------------------------------------------------------------------------------
typedef void (*table_kernel_t)(const TableParams);

template <int NB, int Q, int T>
static table_kernel_t pick_kernel(bool mom, bool full) {
    if (mom && full) return eval_table_kernel<NB, Q, T, true, true>;
    if (mom) return eval_table_kernel<NB, Q, T, true, false>;
    if (full) return eval_table_kernel<NB, Q, T, false, true>;
    return eval_table_kernel<NB, Q, T, false, false>;
}

struct TableConfig {
    table_kernel_t kern = nullptr;
    size_t smem = 0;
    int q = 2, threads = 256, tile = 0;
};

template <int NB, int Q, int T>
static TableConfig make_config(bool mom, bool full, int d) {
    TableConfig c;
    c.q = Q; c.threads = T; c.tile = T * 4 * Q;
    c.smem = tk_smem_bytes<NB, Q, T>(d);
    c.kern = pick_kernel<NB, Q, T>(mom, full);
    return c;
}

template <int NB>
static TableConfig table_config_nb(int d, bool mom, bool full) {
    // shapes: (Q quads/thread, T threads).  Per-draw-sum-only evaluations keep few live registers and take the wide
    // shapes; per-point moments / full output keep 8 points per thread.  BARTINT_TK_SHAPE overrides (tuning).
    int shape = (!mom && !full) ? 1 : 0;                       // 0: q2 t256   1: q4 t256   2: q2 t512
    if (const char* e = getenv("BARTINT_TK_SHAPE")) { if (!mom && !full) shape = atoi(e); }
    if (shape == 1 && tk_smem_bytes<NB, 4, 256>(d) > 112 * 1024) shape = 0;
    if (shape == 2 && tk_smem_bytes<NB, 2, 512>(d) > 112 * 1024) shape = 0;
    switch (shape) {
        case 1: return make_config<NB, 4, 256>(mom, full, d);
        case 2: return make_config<NB, 2, 512>(mom, full, d);
        default: return make_config<NB, 2, 256>(mom, full, d);
    }
}

static TableConfig table_config(int nb, int d, bool mom, bool full) {
    switch (nb) {
        case 4: return table_config_nb<4>(d, mom, full);
        case 5: return table_config_nb<5>(d, mom, full);
        case 6: return table_config_nb<6>(d, mom, full);
        case 7: return table_config_nb<7>(d, mom, full);
        default: return table_config_nb<8>(d, mom, full);
    }
}

// ---------------------------------------------------------------------------------------------
// work decomposition
// ---------------------------------------------------------------------------------------------
EvalPlan make_plan(const bartint_forest* f, const bartint_points* p, uint32_t flags, int32_t draw_begin,
                   int32_t draw_end, bool want_moments, bool want_full) {
    EvalPlan pl{};
    pl.draw_begin = draw_begin;
    pl.draw_end = draw_end;
    const int ndraws = draw_end - draw_begin;
    pl.use_bitslice = f->bs_ok && f->dev.bs_pool != nullptr && !want_moments && !want_full &&
                      !(flags & (BARTINT_FORCE_WALK | BARTINT_NO_BITSLICE));
    if (pl.use_bitslice) {
        // CTA = one tile of points x one permutation block of draws (f->bs_perm) x a share of that block's work units.  Every
        // CTA rebuilds the tile's mask table, so the units are only split when there are too few tiles to occupy the SMs.
        pl.tile = bs_tile(f);
        pl.n_tiles = (int)std::max<int64_t>((p->N + pl.tile - 1) / pl.tile, 1);
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, f->device);
        const int n_pblocks = (draw_end - 1) / f->bs_perm - draw_begin / f->bs_perm + 1;     // blocks the draw range touches
        const int units = 5 * std::min((f->S + 31) / 32, f->bs_perm / 32);
        int uc = std::max(1, std::min(units / 16, (n_sm + pl.n_tiles * n_pblocks - 1) / (pl.n_tiles * n_pblocks)));
        pl.chunk_draws = uc;                                                     // unit chunks per permutation block
        pl.n_chunks = n_pblocks * uc;
        return pl;
    }
    pl.use_table = f->table_ok && !(flags & BARTINT_FORCE_WALK);
    pl.use_gather = pl.use_table && f->gather && p->codes_pm != nullptr;
    pl.count_mode = pl.use_gather && f->count_mode && !want_moments && !want_full;
    int resident_per_sm = 8;
    pl.tile = WK_THREADS;
    if (pl.use_gather) {
        pl.tile = gather_tile(want_moments, want_full);
        resident_per_sm = gather_resident_ctas(f, want_moments, want_full);
    } else if (pl.use_table) {
        TableConfig c = table_config(f->nb, p->d, want_moments, want_full);
        pl.tile = c.tile;
        cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        int nblk = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, c.kern, c.threads, c.smem) != cudaSuccess || nblk < 1) {
            (void)cudaGetLastError();
            nblk = 1;
        }
        resident_per_sm = nblk;
    }
    pl.n_tiles = (int)((p->N + pl.tile - 1) / pl.tile);
    if (pl.n_tiles < 1) pl.n_tiles = 1;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, f->device);
    // Split the draw range into chunks so that the grid is a whole number of full waves of resident CTAs (the tail
    // wave then costs little).  Chunks only add tiny second-stage merges.
    const int resident = n_sm * resident_per_sm;
    int best_chunks = 1;
    double best_eff = -1.0;
    // (the walk kernel with a handful of tiles -- the candidates of a design step -- needs the draws for parallelism)
    const int max_chunks = std::max(1, std::min(ndraws, pl.use_table || pl.n_tiles <= 16 ? 96 : 16));
    for (int c = 1; c <= max_chunks; ++c) {
        const int cd = (ndraws + c - 1) / c;
        const int real_c = (ndraws + cd - 1) / cd;
        if (real_c != c) continue;
        if (cd < 4 && c > 1) continue;
        const double waves = (double)pl.n_tiles * c / resident;
        // each extra chunk re-loads the tile's codes and re-fills the pipeline: mild preference for few chunks
        // (a grid smaller than one wave: every extra chunk simply puts more SMs to work)
        const double eff = waves <= 1.0 ? waves : waves / std::ceil(waves) - 0.002 * c;
        if (eff > best_eff + 1e-9) { best_eff = eff; best_chunks = c; }
    }
    pl.n_chunks = best_chunks;
    pl.chunk_draws = std::max(1, (ndraws + best_chunks - 1) / best_chunks);
    return pl;
}

int launch_eval_table(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                      double* pt_mean, double* pt_m2, double* full_pred, double fp_scale, double fp_y0, int fp_scaled,
                      cudaStream_t st) {
    TableParams prm{};
    prm.group_rec = f->dev.group_rec;
    prm.draw_group_off = f->dev.draw_group_off;
    prm.codes = p->codes;
    prm.nodes = f->dev.nodes;
    prm.tree_off = f->dev.tree_off;
    prm.tree_leaf_off = f->dev.tree_leaf_off;
    prm.leaf_mu = f->dev.leaf_mu;
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
    const TableConfig c = table_config(f->nb, p->d, pt_mean != nullptr, full_pred != nullptr);
    BI_REQUIRE(c.tile == plan.tile, BARTINT_ERR_ARG, "internal: plan/launch tile mismatch");
    BI_CUDA(cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.n_chunks);
    c.kern<<<grid, c.threads, c.smem, st>>>(prm);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// Leaf ordinals decoded from the table kernel's predicate bytes (bit-exactness check of the fast
// path's compare).  One thread per (point quad, group): recompute the predicate bytes with pred_step
// from the same code words and descriptors, then walk each tree of the group with the predicate bits.
// ---------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) table_leaf_kernel(
    const uint8_t* __restrict__ group_rec, const int32_t* __restrict__ draw_group_off,
    const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off, const uint8_t* __restrict__ node_bit,
    const int32_t* __restrict__ tree_perm, const uint8_t* __restrict__ codes, int64_t Npad, int64_t N, int m,
    int draw_begin, int draw_end, int32_t* __restrict__ leaf_out) {
    const int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = draw_group_off[draw_begin] + blockIdx.y;
    if (g >= draw_group_off[draw_end] || quad * 4 >= N) return;
    const uint8_t* rec = group_rec + (int64_t)g * tk_rec_bytes(NB);
    const uint4 h = *reinterpret_cast<const uint4*>(rec);
    if (h.x & TK_FLAG_WALK) {
        const uint2* wn = nodes + tree_off[h.y];
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
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const uint2 dj = desc[j];
        const uint32_t c4 = *reinterpret_cast<const uint32_t*>(codes + (int64_t)dj.x * Npad + quad * 4);
        A = pred_step<NB>(A, c4, dj.y, j);
    }
    for (int kk = 0; kk < 4; ++kk) {
        const int64_t i = quad * 4 + kk;
        if (i >= N) break;
        const uint32_t idx = kk == 0 ? table_index<NB>(A, 0) : kk == 1 ? table_index<NB>(A, 1)
                           : kk == 2 ? table_index<NB>(A, 2) : table_index<NB>(A, 3);
        for (uint32_t t = 0; t < h.z; ++t) {
            const int64_t k = tree_perm[(int64_t)h.y + t];
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
#define LEAF_CASE(NBV)                                                                                              \
    case NBV:                                                                                                       \
        table_leaf_kernel<NBV><<<grid, 256, 0, st>>>(f->dev.group_rec, f->dev.draw_group_off, f->dev.nodes,         \
                                                     f->dev.tree_off, f->dev.node_bit, f->dev.tree_perm, p->codes, \
                                                     p->Npad, p->N, f->m, draw_begin, draw_end, leaf_out);          \
        break;
    switch (f->nb) {
        LEAF_CASE(4) LEAF_CASE(5) LEAF_CASE(6) LEAF_CASE(7) LEAF_CASE(8)
        default: set_error("unsupported table bits %d", f->nb); return BARTINT_ERR_UNSUPPORTED;
    }
#undef LEAF_CASE
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
