// Device-side helpers shared by the evaluation kernels: mbarrier + 1-D bulk (TMA) copies, warp reduction, the
// record-stream pipeline constants and the kernel parameter block.
#pragma once
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

constexpr int TK_NSTAGE = 3;                  // pipeline stages in flight
constexpr int TK_STAGE_TARGET = 12 * 1024;    // bytes of group records per stage (one bulk copy)
__host__ __device__ constexpr int groups_per_stage(int rec_bytes) {   // 4..64 (the per-stage flag masks are 64-bit)
    return TK_STAGE_TARGET / rec_bytes < 4 ? 4 : (TK_STAGE_TARGET / rec_bytes > 64 ? 64 : TK_STAGE_TARGET / rec_bytes);
}

struct TableParams {
    const uint8_t* group_rec;
    const int32_t* draw_group_off;
    const uint8_t* codes;          // feature-major code rows (table kernel; WALK records)
    const uint4* codes_pm;         // point-major packed codes, 16 B per point (gather kernel)
    // global walk arrays for the rare WALK records (trees that do not fit a group)
    const uint2* nodes;
    const int32_t* tree_off;
    const int32_t* tree_leaf_off;
    const double* leaf_mu;
    int64_t Npad, N;
    int d;
    int draw_begin, draw_end, chunk_draws;
    double* draw_part;
    double* pt_mean;
    double* pt_m2;
    double* full_pred;
    double fp_scale, fp_y0;
    int fp_scaled;
    int count_mode;                // 1: skip COUNTABLE records, close draws at END_EVAL (sums-only evaluation)
};

}  // namespace bartint
