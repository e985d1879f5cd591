// Warp-level helpers shared by the fit kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amx {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LC = 32;                    // active-set capacity of the warp solvers (one lane per active atom)
constexpr int TRI = LC * (LC + 1) / 2;    // packed triangular storage
// Active sets that would outgrow the warp solvers are finished by the scalar slow path (k_slow_*).  The limit is a
// constant-bank word so that tests can lower it and exercise that path (AMX_LC_CAP).
__constant__ int c_lc_cap = LC;

__device__ __forceinline__ int tri(int r, int c) { return ((r * (r + 1)) >> 1) + c; }  // r >= c

// 64-bit shuffles as two 32-bit shuffles with plain moves around them: the header's double overload packs its result through
// volatile asm, after which ptxas emitted a three-XOR register swap per use in the substitution loops (5 % of the instructions of
// the first-generation NODDI stage 1, tools/ncu_sass_annot.py).
__device__ __forceinline__ double shfl(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(FULL, lo, src);
    hi = __shfl_sync(FULL, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor(double v, int m)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(FULL, lo, m);
    hi = __shfl_xor_sync(FULL, hi, m);
    return __hiloint2double(hi, lo);
}

// Two interleaved butterfly sums of values that are zero on the lanes >= cnt (warp-uniform cnt): the 32-lane tree
// (offsets 16, 8, 4, 2, 1) adds zeros in its upper levels, so for cnt <= 8 the three lower levels on lanes 0..7 plus one
// broadcast give the bit-identical result for fewer shuffles.  All lanes receive the sums.
__device__ __forceinline__ void warp_sum2_upto(double &a, double &b, int cnt)
{
    if (cnt <= 8) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += shfl_xor(a, o);
            b += shfl_xor(b, o);
        }
        a = shfl(a, 0);
        b = shfl(b, 0);
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += shfl_xor(a, o);
            b += shfl_xor(b, o);
        }
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
    return v;
}

// Order-preserving map double -> uint64 (NaN-free inputs): a < b  <=>  dkey(a) < dkey(b).
__device__ __forceinline__ unsigned long long dkey(double v)
{
    long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

// (value, index) arg-extremum over the lanes with idx >= 0, on the integer reduction unit (REDUX): three
// warp-wide reductions instead of a 5-step shuffle butterfly.  Ties keep the lowest (LOW) or highest index.
// All lanes receive the winner; idx < 0 when no lane had a candidate.
template <bool MAX, bool LOW>
__device__ __forceinline__ void warp_argext(double &v, int &idx)
{
    unsigned long long k = idx >= 0 ? (MAX ? dkey(v) : ~dkey(v)) : 0ull;
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned hm = __reduce_max_sync(FULL, hi);
    unsigned lm = __reduce_max_sync(FULL, hi == hm ? lo : 0u);
    bool win = idx >= 0 && hi == hm && lo == lm;
    int widx = LOW ? __reduce_min_sync(FULL, win ? idx : 0x7fffffff) : __reduce_max_sync(FULL, win ? idx : -1);
    unsigned long long km = ((unsigned long long)hm << 32) | lm;
    bool none = km == 0ull;
    if (!MAX) km = ~km;
    v = dkey_inv(km);
    idx = none ? -1 : widx;
}
// arg-max, ties -> lowest index
__device__ __forceinline__ void warp_argmax(double &v, int &idx) { warp_argext<true, true>(v, idx); }
// arg-min, ties -> lowest (LOWEST) or highest index
template <bool LOWEST>
__device__ __forceinline__ void warp_argmin(double &v, int &idx) { warp_argext<false, LOWEST>(v, idx); }

// arg-min over NON-NEGATIVE values (+inf included): their bit patterns order like unsigned integers, so the order-preserving key
// of warp_argext is not needed.  Lanes with idx < 0 do not take part; ties -> lowest index; idx < 0 when no lane took part.
__device__ __forceinline__ void warp_argmin_nonneg(double &v, int &idx)
{
    const unsigned hi = idx >= 0 ? (unsigned)__double2hiint(v) : 0xffffffffu, lo = idx >= 0 ? (unsigned)__double2loint(v) : 0xffffffffu;
    const unsigned hm = __reduce_min_sync(FULL, hi);
    const unsigned lm = __reduce_min_sync(FULL, hi == hm ? lo : 0xffffffffu);
    const bool win = idx >= 0 && hi == hm && lo == lm;
    const int widx = __reduce_min_sync(FULL, win ? idx : 0x7fffffff);
    v = __hiloint2double((int)hm, (int)lm);
    idx = widx == 0x7fffffff ? -1 : widx;
}

// streaming loads that do not allocate in L1 (the dictionary slab and the signal pass through once per batch;
// L1 is kept for the Gram rows the solvers re-read)
__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
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

// non-blocking probe of an mbarrier phase (no asm labels: can be inlined any number of times); spin with while (!mbar_try(..))
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

}  // namespace amx
