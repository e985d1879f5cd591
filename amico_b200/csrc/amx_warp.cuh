// Warp-level helpers shared by the fit kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amx {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LC = 32;                    // active-set capacity of the warp solvers (one lane per active atom)
constexpr int TRI = LC * (LC + 1) / 2;    // packed triangular storage

__device__ __forceinline__ int tri(int r, int c) { return ((r * (r + 1)) >> 1) + c; }  // r >= c

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// (value, index) arg-max; ties keep the lowest index.  idx < 0 marks "no candidate".
__device__ __forceinline__ void warp_argmax(double &v, int &idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(FULL, v, o);
        int oi = __shfl_xor_sync(FULL, idx, o);
        bool take = (oi >= 0) && (idx < 0 || ov > v || (ov == v && oi < idx));
        if (take) { v = ov; idx = oi; }
    }
}

// (value, index) arg-min; ties keep the lowest index when lowest_on_tie, else the highest.
template <bool LOWEST>
__device__ __forceinline__ void warp_argmin(double &v, int &idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(FULL, v, o);
        int oi = __shfl_xor_sync(FULL, idx, o);
        bool take = (oi >= 0) && (idx < 0 || ov < v || (ov == v && (LOWEST ? oi < idx : oi > idx)));
        if (take) { v = ov; idx = oi; }
    }
}

// un-fused multiply-add: the reference CPU arithmetic (x86-64, no FMA contraction) rounds the product first
__device__ __forceinline__ double madd(double s, double a, double b) { return __dadd_rn(s, __dmul_rn(a, b)); }

// ---- mbarrier + 1-D bulk (TMA) copy ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "AMX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra AMX_DONE;\n"
        "bra AMX_WAIT;\n"
        "AMX_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP); bytes % 16 == 0, 16-B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace amx
