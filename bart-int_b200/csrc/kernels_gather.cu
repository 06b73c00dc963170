// K1 (gather variant)  posterior ensemble evaluation with the points' codes held in REGISTERS.
//
// Same contract as eval_table_kernel (kernels_eval.cu): group records streamed through shared memory by 1-D bulk
// TMA copies + an mbarrier ring, joint fp64 leaf table per group, fused per-draw sums / per-point moments.  The
// difference is how a group's predicate bits are formed.  The table kernel loads one code word per (node, 4 points)
// from shared memory -- a third of its shared-memory wavefronts, and the shared-memory pipe is what bounds it.
// Here every point's d <= 16 codes live in R = 2..4 registers (point-major, 4 codes per word) and per (group, point):
//     g = PRMT(words, sel)          gather the codes of the (up to 4) nodes of a round      (selector: per group)
//     m = PRMT.sign(g + bias)       0xFF per byte iff code > cut                           (bias = 127 - cut)
//     o = IDP.4A(m, w, o)           w = -(8 << bit) per slot  ->  o = table byte offset idx * 8
// twice (two rounds: variables 0..7 and 4(R-2)..4(R-2)+7), then one LDS.64 of the joint table and one DADD.
// Branch-free, warp-uniform, and the only shared-memory traffic left is the table lookup + 2 descriptor loads.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "pipeline.cuh"

namespace bartint {

constexpr int TG_THREADS = 256;
constexpr int TG_P = 8;                          // points per thread (16 in the sums-only kernel, see gather_tile)

// ---------------------------------------------------------------------------------------------
// point-major packed codes: codes_pm[i] = 16 bytes, byte j = code of variable j (0 beyond d)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) repack_codes_kernel(const uint8_t* __restrict__ codes, int64_t Npad, int d,
                                                           uint4* __restrict__ codes_pm, int64_t rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int j = 0; j < d; ++j) w[j >> 2] |= (uint32_t)codes[(int64_t)j * Npad + i] << (8 * (j & 3));
    codes_pm[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

int launch_repack_codes(const uint8_t* codes, int64_t Npad, int d, uint4* codes_pm, cudaStream_t st, int64_t rows) {
    if (Npad == 0) return BARTINT_OK;
    if (rows <= 0) rows = Npad;          // rows > 0: a row block (codes / codes_pm point at its first row; stride Npad)
    repack_codes_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(codes, Npad, d, codes_pm, rows);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// the gather step, SHARED by the evaluation kernel and the leaf-decoding kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_mask4(uint32_t v) {
    uint32_t m;   // per byte: 0xFF if bit 7 set else 0x00 (selector nibbles 8..B replicate the sign; PTX only)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(v), "r"(0u), "r"(0xBA98u));
    return m;
}

// desc words: d0 = selA | selB << 16, d1 = biasA, d2 = biasB, d3 = wA, d4 = wB.   Returns idx * 8.
template <int R>
__device__ __forceinline__ uint32_t gather_offset(const uint32_t (&c)[R], uint32_t d0, uint32_t d1, uint32_t d2,
                                                  uint32_t d3, uint32_t d4) {
    const uint32_t gA = __byte_perm(c[0], c[1], d0);                 // nibbles 0..3 of d0
    const uint32_t gB = __byte_perm(c[R - 2], c[R - 1], d0 >> 16);   // nibbles 4..7 of d0
    int o = __dp4a((int)sign_mask4(gA + d1), (int)d3, 0);
    o = __dp4a((int)sign_mask4(gB + d2), (int)d4, o);
    return (uint32_t)o;
}

// DUAL record: two independent 16-entry tables.  Table 1 gathers from code words (0,1); table 2 from (R-2,R-1)
// [V = 0], (0,R-1) [V = 1] or (PRMT(w0,w1,sel_t), R-1) [V = 2: any three words, one more PRMT].  V is a property of
// the record (warp-uniform branch outside the point loop).  Returns both byte offsets (table 2's includes its +128).
template <int R, int V>
__device__ __forceinline__ void dual_offsets(const uint32_t (&c)[R], uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3,
                                             uint32_t d4, uint32_t sel_t, uint32_t& o1, uint32_t& o2) {
    const uint32_t lo = (R == 2 || V == 1) ? c[0] : (V == 0) ? c[R - 2] : __byte_perm(c[0], c[1], sel_t);
    const uint32_t gA = __byte_perm(c[0], c[1], d0);
    const uint32_t gB = __byte_perm(lo, c[R - 1], d0 >> 16);
    o1 = (uint32_t)__dp4a((int)sign_mask4(gA + d1), (int)d3, 0);
    o2 = (uint32_t)__dp4a((int)sign_mask4(gB + d2), (int)d4, 128);
}

template <int R>
__device__ __forceinline__ uint32_t code_of(const uint32_t (&c)[R], uint32_t var) {   // WALK records: one code by variable
    uint32_t lo = c[0], hi = c[1];
    if (R > 2 && var >= 8) { lo = c[R - 2]; hi = c[R - 1]; var -= 4 * (R - 2); }
    return __byte_perm(lo, hi, 0x4440u | var) & 0xFFu;
}

template <int NB>
size_t tg_smem_bytes() {
    constexpr int GS = groups_per_stage(tg_rec_bytes(NB));
    return (size_t)TK_NSTAGE * GS * tg_rec_bytes(NB) + 2 * GS * (TG_THREADS / 32) * sizeof(double) +
           (TK_NSTAGE + 1) * sizeof(uint64_t);
}

// COUNT = true is the sums-only evaluation of a counting-mode forest (skip COUNTABLE records, close draws at END_EVAL).
// It is a compile-time variant so that the default kernel carries no trace of it: with the mode tested at run time
// inside the stage loop the 16-point body spilled (STACK 32 -> 48, round-1 regression 5.56 -> 6.18 ms at cfg 2).
template <int NB, int R, bool MOMENTS, bool FULL, int P = TG_P, bool COUNT = false>
__global__ void __launch_bounds__(TG_THREADS, P > 8 ? 2 : (MOMENTS || FULL) ? 2 : 3) eval_gather_kernel(const TableParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = TG_THREADS, TILE = TG_THREADS * P;
    constexpr int REC = tg_rec_bytes(NB);
    constexpr int GS = groups_per_stage(REC);
    constexpr int STAGE = GS * REC;
    constexpr int NWARP = T / 32;
    unsigned char* s_stage = smem;                                             // TK_NSTAGE stage buffers
    double* s_wsum = reinterpret_cast<double*>(s_stage + TK_NSTAGE * STAGE);   // [2][GS][NWARP]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_wsum + 2 * GS * NWARP);    // [TK_NSTAGE]

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
        for (int b = 0; b < TK_NSTAGE; ++b) mbar_init(&s_bar[b], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < TK_NSTAGE && k < n_stages; ++k) {
            const int ng = min(GS, n_groups - k * GS);
            mbar_expect_tx(&s_bar[k], (uint32_t)(ng * REC));
            bulk_g2s(s_stage + k * STAGE, rec_g + (int64_t)k * STAGE, (uint32_t)(ng * REC), &s_bar[k]);
        }
    }

    // this thread's points: tile point p * T + tid (warp-contiguous 16 B loads); codes stay in registers
    const int64_t tile_base = (int64_t)tile * TILE;
    uint32_t creg[P][R];
    uint32_t vmask = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int64_t i = tile_base + (int64_t)p * T + tid;            // < Npad (zero padded), always readable
        const uint4 v = __ldg(&prm.codes_pm[i]);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < R; ++r) creg[p][r] = w[r];
        if (i < prm.N) vmask |= 1u << p;
    }

    double acc[P];
    double piv[MOMENTS ? P : 1], S1[MOMENTS ? P : 1], S2[MOMENTS ? P : 1];
    // FULL (the S x N matrix of predict(), draws contiguous per point in R's layout): the draws of the current group of
    // four rows wait in registers (slot = row mod 4, statically indexed) and leave as ONE aligned 32-byte sector per
    // point -- two 16-byte stores -- whenever the draw index reaches a multiple of four; a warp that stored one
    // double per draw touched 32 sectors for 256 bytes of payload.  Ragged heads / tails of a draw chunk and row
    // counts that are not a multiple of four fall back to scalar stores.
    double fq[FULL ? P : 1][4];
    int fq_n = 0;                                              // values waiting (uniform across the CTA)
    const bool fq_vec = FULL && (ndraws & 3) == 0 && (reinterpret_cast<uintptr_t>(prm.full_pred) & 31) == 0;
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.0;
#pragma unroll
    for (int p = 0; p < (MOMENTS ? P : 1); ++p) { piv[p] = 0.0; S1[p] = 0.0; S2[p] = 0.0; }

    int draws_done = 0;   // uniform across the CTA
    for (int k = 0; k < n_stages; ++k) {
        const int b = k % TK_NSTAGE;
        mbar_wait(&s_bar[b], (uint32_t)((k / TK_NSTAGE) & 1));
        const int ng = min(GS, n_groups - k * GS);
        const unsigned char* stage = s_stage + b * STAGE;
        const int par = k & 1;
        const int stage_first_draw = draws_done;
        // the stage's group flags as two warp-uniform bit masks
        unsigned long long walk_mask = 0ull, end_mask = 0ull, joint_mask = 0ull, skip_mask = 0ull;
#pragma unroll
        for (int base = 0; base < GS; base += 32) {
            const int gl = base + lane;
            const uint32_t fl = gl < ng ? *reinterpret_cast<const uint32_t*>(stage + gl * REC) : 0u;
            constexpr uint32_t end_flag = COUNT ? TK_FLAG_END_EVAL : TK_FLAG_END_DRAW;
            walk_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & TK_FLAG_WALK) != 0u) << base;
            end_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & end_flag) != 0u) << base;
            joint_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & TK_FLAG_JOINT) != 0u) << base;
            if (COUNT) skip_mask |= (unsigned long long)__ballot_sync(0xffffffffu, (fl & TK_FLAG_COUNTABLE) != 0u) << base;
        }
        // DUAL records without a flag (the bulk) run in a tight loop; the next flagged record is found with one ffs
        auto dual_points = [&](auto variant, const uint4& da, uint32_t wB, uint32_t sel_t,
                               const unsigned char* __restrict__ tbl) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                uint32_t o1, o2;
                dual_offsets<R, decltype(variant)::value>(creg[p], da.x, da.y, da.z, da.w, wB, sel_t, o1, o2);
                acc[p] += *reinterpret_cast<const double*>(tbl + o1);
                acc[p] += *reinterpret_cast<const double*>(tbl + o2);
            }
        };
        auto dual_group = [&](const unsigned char* rec) {
            // smem broadcast loads: 16 + 8 bytes (a broadcast LDS.128 costs 2 wavefronts, an LDS.64 one); sel_t is
            // fetched only by the three-word kind
            const uint4 da = *reinterpret_cast<const uint4*>(rec + TK_HDR);
            const uint2 db = *reinterpret_cast<const uint2*>(rec + TK_HDR + 16);       // wB, table-2 kind
            const unsigned char* __restrict__ tbl = rec + tg_tbl_off();
            if (R == 2 || db.y == 0u) dual_points(std::integral_constant<int, 0>{}, da, db.x, 0u, tbl);
            else if (db.y == 1u) dual_points(std::integral_constant<int, 1>{}, da, db.x, 0u, tbl);
            else dual_points(std::integral_constant<int, 2>{}, da, db.x, *reinterpret_cast<const uint32_t*>(rec + TK_HDR + 24), tbl);
        };
        auto joint_group = [&](const unsigned char* rec) {
            const uint4 da = *reinterpret_cast<const uint4*>(rec + TK_HDR);
            const uint32_t d4 = *reinterpret_cast<const uint32_t*>(rec + TK_HDR + 16);
            const unsigned char* __restrict__ tbl = rec + tg_tbl_off();
#pragma unroll
            for (int p = 0; p < P; ++p)
                acc[p] += *reinterpret_cast<const double*>(tbl + gather_offset<R>(creg[p], da.x, da.y, da.z, da.w, d4));
        };
        int g = 0;
#pragma unroll 1
        while (g < ng) {
            const unsigned long long special = (COUNT ? (walk_mask | end_mask | joint_mask | skip_mask) : (walk_mask | end_mask | joint_mask)) >> g;
            const int g_stop = special ? min(ng, g + __ffsll((long long)special) - 1) : ng;
#pragma unroll 1
            for (; g < g_stop; ++g) dual_group(stage + g * REC);
            if (g >= ng) break;
            const unsigned char* rec = stage + g * REC;
            if (COUNT && ((skip_mask >> g) & 1ull)) { ++g; continue; }      // this record's trees are counted, not evaluated
            if ((walk_mask >> g) & 1ull) {
                // rare: a tree that does not fit a group, walked per point from global memory (L1/L2)
                const int64_t kt = reinterpret_cast<const uint32_t*>(rec)[1];
                const uint2* __restrict__ wn = prm.nodes + prm.tree_off[kt];
                const double* __restrict__ wmu = prm.leaf_mu + prm.tree_leaf_off[kt];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint32_t node = 0;
                    uint2 r = wn[0];
                    while ((r.x & 0xFFFFu) != LEAF_VAR) {
                        node += (code_of<R>(creg[p], r.x & 0xFFFFu) > (r.x >> 16)) ? r.y : 1u;
                        r = wn[node];
                    }
                    acc[p] += wmu[r.y];
                }
            } else if ((joint_mask >> g) & 1ull) {
                joint_group(rec);
            } else {
                dual_group(rec);
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
                    if (FULL) {                                       // slot = row mod 4 (warp-uniform, static indices)
                        const double fo = prm.fp_scaled ? fv : prm.fp_scale * (fv + 0.5) + prm.fp_y0;
                        const int slot = (s - prm.draw_begin) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j) fq[p][j] = slot == j ? fo : fq[p][j];     // selects: fq stays in registers
                    }
                }
                if (FULL) {
                    ++fq_n;
                    const int row = s - prm.draw_begin, slot = row & 3;       // row of the newest value
                    if (slot == 3 || s + 1 == s1) {                           // a sector is complete, or the chunk ends
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            if ((vmask >> p) & 1u) {
                                double* dst = prm.full_pred + (int64_t)ndraws * (tile_base + (int64_t)p * T + tid) + (row - slot);
                                if (fq_n == 4 && fq_vec) {                    // (fq_n == 4 implies slot == 3)
                                    *reinterpret_cast<double2*>(dst) = make_double2(fq[p][0], fq[p][1]);
                                    *reinterpret_cast<double2*>(dst + 2) = make_double2(fq[p][2], fq[p][3]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (j <= slot && j > slot - fq_n) dst[j] = fq[p][j];
                                }
                            }
                        }
                        fq_n = 0;
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
                const int64_t i = tile_base + (int64_t)p * T + tid;
                prm.pt_mean[(int64_t)chunk * prm.N + i] = piv[p] + S1[p] / n;
                prm.pt_m2[(int64_t)chunk * prm.N + i] = S2[p] - S1[p] * S1[p] / n;
            }
        }
    }
}

// points per thread: 16 for the sums-only kernel (the Monte Carlo path: 2 CTAs x 256 threads x 128 registers per SM;
// halves the per-record descriptor/loop overhead per point, 5.76 -> 5.56 ms at cfg 2), 8 when per-point moments or
// the full matrix are wanted (their extra accumulators do not fit 16 points).  BARTINT_GATHER_P=8 forces 8.
int gather_points_per_thread(bool mom, bool full) {
    static const bool p8 = [] { const char* e = std::getenv("BARTINT_GATHER_P"); return e && std::atoi(e) == 8; }();
    if (mom && full) return 4;        // moments + the four-row store buffer: 8 points would spill 200 bytes per thread
    return (!p8 && !mom && !full) ? 16 : TG_P;
}
int gather_tile(bool mom, bool full) { return TG_THREADS * gather_points_per_thread(mom, full); }

// ---- instantiation table -----------------------------------------------------------------------------------
typedef void (*gather_kernel_t)(const TableParams);

template <int NB, int R>
static gather_kernel_t pick_gather(bool mom, bool full, bool count) {
    if (mom && full) return eval_gather_kernel<NB, R, true, true, 4>;
    if (mom) return eval_gather_kernel<NB, R, true, false>;
    if (full) return eval_gather_kernel<NB, R, false, true>;
    if (gather_points_per_thread(mom, full) == 16)
        return count ? eval_gather_kernel<NB, R, false, false, 16, true> : eval_gather_kernel<NB, R, false, false, 16>;
    return count ? eval_gather_kernel<NB, R, false, false, TG_P, true> : eval_gather_kernel<NB, R, false, false>;
}

template <int NB>
static gather_kernel_t pick_gather_r(int regs, bool mom, bool full, bool count) {
    switch (regs) {
        case 2: return pick_gather<NB, 2>(mom, full, count);
        case 3: return pick_gather<NB, 3>(mom, full, count);
        default: return pick_gather<NB, 4>(mom, full, count);
    }
}

// count: the sums-only evaluation of a counting-mode forest (EvalPlan::count_mode); never with mom / full
static gather_kernel_t gather_kernel_for(const bartint_forest* f, bool mom, bool full, bool count, size_t* smem) {
    if (f->nb == 4) { *smem = tg_smem_bytes<4>(); return pick_gather_r<4>(f->g_regs, mom, full, count); }
    *smem = tg_smem_bytes<5>();
    return pick_gather_r<5>(f->g_regs, mom, full, count);
}

int gather_resident_ctas(const bartint_forest* f, bool mom, bool full) {
    size_t smem = 0;
    gather_kernel_t k = gather_kernel_for(f, mom, full, f->count_mode && !mom && !full, &smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nblk = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, k, TG_THREADS, smem) != cudaSuccess || nblk < 1) {
        (void)cudaGetLastError();
        nblk = 1;
    }
    return nblk;
}

int launch_eval_gather(const bartint_forest* f, const bartint_points* p, const EvalPlan& plan, double* draw_part,
                       double* pt_mean, double* pt_m2, double* full_pred, double fp_scale, double fp_y0, int fp_scaled,
                       cudaStream_t st) {
    BI_REQUIRE(p->codes_pm != nullptr, BARTINT_ERR_ARG, "internal: points have no packed codes for the gather kernel");
    BI_REQUIRE(plan.tile == gather_tile(pt_mean != nullptr, full_pred != nullptr), BARTINT_ERR_ARG,
               "internal: plan/launch tile mismatch");
    TableParams prm{};
    prm.group_rec = f->dev.group_rec;
    prm.draw_group_off = f->dev.draw_group_off;
    prm.codes = p->codes;
    prm.codes_pm = p->codes_pm;
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
    prm.count_mode = plan.count_mode ? 1 : 0;
    size_t smem = 0;
    gather_kernel_t k = gather_kernel_for(f, pt_mean != nullptr, full_pred != nullptr, plan.count_mode, &smem);
    BI_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.n_chunks);
    k<<<grid, TG_THREADS, smem, st>>>(prm);
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

// ---------------------------------------------------------------------------------------------
// Leaf ordinals decoded from the gather kernel's table index (bit-exactness check of the fast path's compare):
// one thread per (point, group) recomputes gather_offset from the same packed codes and descriptors, then walks each
// tree of the group with the predicate bits.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) gather_leaf_kernel(
    const uint8_t* __restrict__ group_rec, int nb, const int32_t* __restrict__ draw_group_off,
    const uint2* __restrict__ nodes, const int32_t* __restrict__ tree_off, const uint8_t* __restrict__ node_bit,
    const int32_t* __restrict__ tree_perm, const uint4* __restrict__ codes_pm, int64_t N, int m, int draw_begin,
    int draw_end, int32_t* __restrict__ leaf_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = draw_group_off[draw_begin] + blockIdx.y;
    if (g >= draw_group_off[draw_end] || i >= N) return;
    const uint8_t* rec = group_rec + (int64_t)g * tg_rec_bytes(nb);
    const uint4 h = *reinterpret_cast<const uint4*>(rec);
    const uint4 v = codes_pm[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = w[r];
    if (h.x & TK_FLAG_WALK) {
        const uint2* wn = nodes + tree_off[h.y];
        uint32_t node = 0;
        uint2 r = wn[0];
        while ((r.x & 0xFFFFu) != LEAF_VAR) {
            node += (code_of<R>(c, r.x & 0xFFFFu) > (r.x >> 16)) ? r.y : 1u;
            r = wn[node];
        }
        leaf_out[((int64_t)h.y - (int64_t)draw_begin * m) * N + i] = (int32_t)r.y;
        return;
    }
    const uint32_t* dw = reinterpret_cast<const uint32_t*>(rec + TK_HDR);
    uint32_t idx1, idx2;
    if (h.x & TK_FLAG_JOINT) {
        idx1 = idx2 = gather_offset<R>(c, dw[0], dw[1], dw[2], dw[3], dw[4]) >> 3;
    } else {
        uint32_t o1, o2;
        if (R == 2 || dw[5] == 0u) dual_offsets<R, 0>(c, dw[0], dw[1], dw[2], dw[3], dw[4], dw[6], o1, o2);
        else if (dw[5] == 1u) dual_offsets<R, 1>(c, dw[0], dw[1], dw[2], dw[3], dw[4], dw[6], o1, o2);
        else dual_offsets<R, 2>(c, dw[0], dw[1], dw[2], dw[3], dw[4], dw[6], o1, o2);
        idx1 = o1 >> 3;
        idx2 = (o2 - 128u) >> 3;
    }
    for (uint32_t t = 0; t < h.z; ++t) {
        const uint32_t idx = (h.x & TK_FLAG_JOINT) || t < h.w ? idx1 : idx2;
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

int launch_gather_leaf(const bartint_forest* f, const bartint_points* p, int32_t draw_begin, int32_t draw_end,
                       int32_t* leaf_out, cudaStream_t st) {
    const int n_groups = f->draw_group_off[(size_t)draw_end] - f->draw_group_off[(size_t)draw_begin];
    if (n_groups == 0 || p->N == 0) return BARTINT_OK;
    BI_REQUIRE(p->codes_pm != nullptr, BARTINT_ERR_ARG, "internal: points have no packed codes for the gather kernel");
    BI_REQUIRE(n_groups <= 65535, BARTINT_ERR_UNSUPPORTED, "leaf decoding limited to 65535 groups per call");
    dim3 grid((unsigned)((p->N + 255) / 256), (unsigned)n_groups);
#define GL_CASE(RV)                                                                                                \
    case RV:                                                                                                       \
        gather_leaf_kernel<RV><<<grid, 256, 0, st>>>(f->dev.group_rec, f->nb, f->dev.draw_group_off, f->dev.nodes, \
                                                     f->dev.tree_off, f->dev.node_bit, f->dev.tree_perm, p->codes_pm, p->N, \
                                                     f->m, draw_begin, draw_end, leaf_out);                        \
        break;
    switch (f->g_regs) {
        GL_CASE(2) GL_CASE(3)
        default: gather_leaf_kernel<4><<<grid, 256, 0, st>>>(f->dev.group_rec, f->nb, f->dev.draw_group_off, f->dev.nodes,
                                                             f->dev.tree_off, f->dev.node_bit, f->dev.tree_perm, p->codes_pm,
                                                             p->N, f->m, draw_begin, draw_end, leaf_out);
    }
#undef GL_CASE
    count_launch();
    BI_CUDA(cudaGetLastError());
    return BARTINT_OK;
}

}  // namespace bartint
