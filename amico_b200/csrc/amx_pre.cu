// The callers either side of the per-voxel fit, on the GPU (include/amico_b200.h, second part):
//   amx_volume_to_voxel_major  nibabel get_fdata().astype(float32) + voxel-major layout, amico/core.py:135-136
//   amx_preprocess             amico/core.py:151-156, 209-278 (load_data: NaN policy, b0 normalisation, b0 merge, directional
//                              average) + :451-452 (mask compaction, y < 0 -> 0)
//   amx_mean_b0                amico/core.py:212
//   amx_dti_directions         amico/core.py:430-436, 456-458 (dipy TensorModel OLS -> principal eigenvector)
//   amx_scatter_maps           amico/core.py:472-498
//   amx_resample_kernels       amico/lut.pyx:274-311 (+ the [:, merge_idx] of every <Model>.resample)
//
// All are streaming passes over voxel-major data (a few flops per byte, fp64 log chains for the DTI fit).  k_preprocess: a warp
// takes a CHUNK of 32 consecutive voxels -- one contiguous block of 32 * nS floats -- as ONE TMA bulk copy into shared memory
// (double buffered), lane v does voxel v's short sequential arithmetic, and the rows stream back out as float4.  Grids are
// sized to the SM count and walk the chunks with a stride (persistent warps).
#include "../../include/amico_b200.h"
#include "amx_err.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int CH = 32;           // voxels per chunk (one per lane)
constexpr int PRE_BLOCK_VOX = 1024;  // voxels per mask-count block

struct PreParams {
    const float *dwi; long long n_total; int nS;
    const unsigned char *mask;
    const int *b0_idx; int b0_count;
    const int *dwi_idx; int dwi_count;
    const int *shell_idx; const int *shell_off; int n_shells;
    unsigned flags; float thr, repl;
    int m_out;
    float *y; long long y_cap; int *vox_idx; float *mean_b0s;
    const long long *block_off;  // exclusive prefix of kept voxels per PRE_BLOCK_VOX block
    unsigned *status;            // [0] bit0: non-finite raw value, bit1: non-finite pre-processed value, bit2: y overflow
};

// ---- mask compaction: counts per block of 1024 voxels, exclusive scan, total -------------------------------------
__global__ void k_mask_count(const unsigned char *__restrict__ mask, long long n, long long *counts)
{
    const long long base = (long long)blockIdx.x * PRE_BLOCK_VOX;
    int c = 0;
    for (int i = threadIdx.x; i < PRE_BLOCK_VOX; i += blockDim.x) {
        const long long v = base + i;
        if (v < n) c += mask ? (mask[v] == 1) : 1;
    }
    __shared__ int ws[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULLM, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLM, t, o);
        if (threadIdx.x == 0) counts[blockIdx.x] = t;
    }
}

// single-block exclusive scan over n_blocks counts (in place), total -> counts[n_blocks]
__global__ void k_scan_counts(long long *counts, int n_blocks)
{
    __shared__ long long wtot[32], wexcl[32];
    __shared__ long long carry, tot;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b0 = 0; b0 < n_blocks; b0 += blockDim.x) {
        const int i = b0 + threadIdx.x;
        const long long v = i < n_blocks ? counts[i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(FULLM, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const long long w = lane < nw ? wtot[lane] : 0;
            long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(FULLM, winc, o);
                if (lane >= o) winc += t;
            }
            wexcl[lane] = winc - w;
            if (lane == 31) tot = winc;
        }
        __syncthreads();
        if (i < n_blocks) counts[i] = carry + wexcl[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n_blocks] = carry;
}

// ---- shared-memory staging of a chunk ------------------------------------------------------------------------------
// count = nvox * nS consecutive floats starting at src (16-byte aligned: chunks start at multiples of 32 voxels) ->
// rows of `stride` words.  Returns true when a non-finite value passed through (raw-signal check of core.py:151).
__device__ __forceinline__ bool stage_rows(float *rows, const float *__restrict__ src, int count, int nS, int stride, int lane,
                                           bool replace, float repl)
{
    bool bad = false;
    int e = lane * 4;
    int v = e / nS, j = e - v * nS;
    const int count4 = count & ~3;
    while (e < count4) {
        float4 q = __ldcs(reinterpret_cast<const float4 *>(src + e));
        float x[4] = {q.x, q.y, q.z, q.w};
        int vv = v, jj = j;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float f = x[t];
            if (!isfinite(f)) {
                bad = true;
                if (replace) f = repl;
            }
            rows[vv * stride + jj] = f;
            if (++jj == nS) { jj = 0; ++vv; }
        }
        e += 128;
        j += 128;
        while (j >= nS) { j -= nS; ++v; }
    }
    // tail (count not a multiple of 4 can only happen in the last, partial chunk)
    for (int t = count4 + lane; t < count; t += 32) {
        float f = __ldcs(src + t);
        if (!isfinite(f)) {
            bad = true;
            if (replace) f = repl;
        }
        const int vv = t / nS;
        rows[vv * stride + (t - vv * nS)] = f;
    }
    return bad;
}

// sequential float32 mean, index order: what numpy does for np.mean(img[:, :, :, idx], axis=3) (the fancy-indexed copy
// has the indexed axis slowest in memory, so the reduction is a plain in-order accumulation, then one division)
template <typename F>
__device__ __forceinline__ float seq_mean(int n, F get)
{
    float s = get(0);
    for (int i = 1; i < n; ++i) s = __fadd_rn(s, get(i));
    return __fdiv_rn(s, (float)n);
}

// ---- mbarrier + 1-D bulk (TMA) copy, as in amx_warp.cuh ------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "PRE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra PRE_DONE;\n"
        "bra PRE_WAIT;\n"
        "PRE_DONE:\n"
        "}\n" ::"r"(s_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
                 "l"(src), "r"(bytes), "r"(s_u32(bar))
                 : "memory");
}

// bytes of shared memory one warp of k_preprocess needs
__host__ __device__ inline size_t pre_warp_bytes(int nS, int m_out, bool diravg)
{
    size_t b = 2 * (size_t)CH * nS * sizeof(float);                 // two raw chunk buffers (TMA destinations)
    b += 2 * CH * sizeof(float);                                     // norm factor, merged b0
    if (diravg) b += (size_t)CH * ((nS | 1) + m_out) * sizeof(float);  // padded copy for the lane-per-voxel sums + averages
    return (b + 127) & ~(size_t)127;
}

// One warp per chunk of 32 consecutive voxels = one contiguous block of 32 * nS floats.  Full chunks arrive by ONE bulk
// (TMA) copy each, double buffered: the copy of the warp's next chunk is in flight while it works on the current one.
//   pass 1 (cooperative): NaN / Inf policy on the raw values (core.py:151-156)
//   pass 2 (lane v = voxel v): mean b0, norm factor (core.py:212-219); merged b0 / shell averages when asked
//   pass 3 (cooperative): y rows of the kept voxels, normalised, clamped at 0, written fully coalesced
__global__ void __launch_bounds__(128) k_preprocess(const PreParams p, long long n_chunks, unsigned warp_bytes)
{
    extern __shared__ __align__(128) unsigned char pre_smem[];
    __shared__ uint64_t bars[4][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nS = p.nS, m_out = p.m_out;
    const bool norm = p.flags & AMX_PRE_NORMALIZE, merge = p.flags & AMX_PRE_MERGE_B0, diravg = p.flags & AMX_PRE_DIR_AVG;
    const bool replace = p.flags & AMX_PRE_REPLACE_BAD;
    float *buf0 = reinterpret_cast<float *>(pre_smem + (size_t)warp * warp_bytes);
    float *nfs = buf0 + 2 * CH * nS, *mb0s = nfs + CH;
    const int pstride = nS | 1;
    float *pad = mb0s + CH, *avg = pad + CH * pstride;  // diravg only
    uint64_t *bar = bars[warp];
    if (lane == 0) {
        mb_init(&bar[0], 1);
        mb_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long step = (long long)gridDim.x * wpb;
    const uint32_t chunk_bytes = (uint32_t)(CH * nS * sizeof(float));
    unsigned st = 0;
    long long c = (long long)blockIdx.x * wpb + warp;
    auto load = [&](long long cc, int b) {  // all lanes call; full chunks: one TMA copy, the partial last chunk: plain loads
        const long long v0 = cc * CH;
        const int nvox = (int)min((long long)CH, p.n_total - v0);
        float *dst = buf0 + (size_t)b * CH * nS;
        if (nvox == CH) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our earlier generic accesses to dst come first
                mb_expect_tx(&bar[b], chunk_bytes);
                tma_g2s(dst, p.dwi + v0 * nS, chunk_bytes, &bar[b]);
            }
        } else {
            const float *src = p.dwi + v0 * nS;
            for (int e = lane; e < nvox * nS; e += 32) dst[e] = __ldcs(src + e);
        }
    };
    if (c < n_chunks) load(c, 0);
    for (int it = 0; c < n_chunks; c += step, ++it) {
        const int b = it & 1;
        const long long v0 = c * CH;
        const int nvox = (int)min((long long)CH, p.n_total - v0);
        __syncwarp();
        if (c + step < n_chunks) load(c + step, b ^ 1);
        if (nvox == CH) mb_wait(&bar[b], (it >> 1) & 1);
        __syncwarp();
        float *raw = buf0 + (size_t)b * CH * nS;
        // mask first: whether pass 1 can be folded into the vectorised pass 3 depends on it
        bool keep = false;
        if (lane < nvox) keep = p.mask ? (p.mask[v0 + lane] == 1) : true;
        const unsigned keepmask = __ballot_sync(FULLM, keep);
        // fast path: every voxel of a full chunk is kept and the output row is the whole input row -> pass 3 streams the
        // chunk as float4 and checks the raw values on the way, no separate pass 1
        const bool fast = !merge && !diravg && !replace && keepmask == FULLM && (nS & 3) == 0 && nvox == CH;
        // ---- pass 1
        if (!fast) {
            bool bad = false;
            const int cnt = nvox * nS;
            if (diravg) {  // also copy into rows of odd stride for the conflict-free lane-per-voxel sums below
                int v = 0, j = lane;
                while (j >= nS) { j -= nS; ++v; }
                for (int e = lane; e < cnt; e += 32) {
                    float f = raw[e];
                    if (!isfinite(f)) { bad = true; if (replace) f = p.repl; }
                    pad[v * pstride + j] = f;
                    j += 32;
                    while (j >= nS) { j -= nS; ++v; }
                }
            } else {
#pragma unroll 4
                for (int e = lane; e < cnt; e += 32) {
                    const float f = raw[e];
                    if (!isfinite(f)) { bad = true; if (replace) raw[e] = p.repl; }
                }
            }
            if (bad) st |= 1u;
        }
        __syncwarp();
        // ---- pass 2
        if (lane < nvox) {
            const float *r = diravg ? pad + lane * pstride : raw + lane * nS;
            float nf = 1.0f;
            if (p.b0_count > 0 && (norm || p.mean_b0s)) {
                const float mb = seq_mean(p.b0_count, [&](int i) { return r[p.b0_idx[i]]; });
                if (p.mean_b0s) p.mean_b0s[v0 + lane] = mb;
                // norm_factor = mean_b0s; idx = nf <= thr; nf[idx] = 1; nf = 1 / nf; nf[idx] = 0   (core.py:215-219)
                if (norm) nf = (mb <= p.thr) ? 0.0f : __fdiv_rn(1.0f, mb);
            }
            nfs[lane] = nf;
            if (merge)  // mean of the (normalised) b0 volumes (core.py:226)
                mb0s[lane] = seq_mean(p.b0_count, [&](int i) { const float x = r[p.b0_idx[i]]; return norm ? __fmul_rn(x, nf) : x; });
            if (diravg) {
                // dir_avg_img is a VIEW of the first n_shells+1 volumes (core.py:234): every mean is written in place, so a
                // later mean that indexes volume i < (volumes written so far) reads the average stored there
                float *a = avg + lane * m_out;
                int done = 0;
                auto val = [&](int i) { return i < done ? a[i] : (norm ? __fmul_rn(r[i], nf) : r[i]); };
                a[0] = seq_mean(p.b0_count, [&](int i) { return val(p.b0_idx[i]); });
                done = 1;
                for (int s = 0; s < p.n_shells; ++s) {
                    const int o = p.shell_off[s], cnt = p.shell_off[s + 1] - o;
                    a[s + 1] = seq_mean(cnt, [&](int i) { return val(p.shell_idx[o + i]); });
                    done = s + 2;
                }
            }
        }
        __syncwarp();
        // ---- pass 3: the kept voxels' rows are consecutive in y
        const int nkeep = __popc(keepmask);
        if (fast) {
            long long pos = v0;
            if (p.mask) {
                const long long b0v = (v0 / PRE_BLOCK_VOX) * PRE_BLOCK_VOX;
                int before = 0;
                for (long long v = b0v + lane; v < v0; v += 32) before += (p.mask[v] == 1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(FULLM, before, o);
                pos = p.block_off[v0 / PRE_BLOCK_VOX] + before;
            }
            if (pos + CH > p.y_cap) {
                st |= 4u;
            } else {
                p.vox_idx[pos + lane] = (int)(v0 + lane);
                const float4 *src4 = reinterpret_cast<const float4 *>(raw);
                float4 *out4 = reinterpret_cast<float4 *>(p.y + pos * nS);
                const int q4 = nS >> 2, total4 = CH * q4;
                int v = 0, j4 = lane;
                while (j4 >= q4) { j4 -= q4; ++v; }
                float chk_raw = 0.0f, chk_out = 0.0f;  // x * 0 is NaN exactly when x is NaN or +-Inf
#pragma unroll 2
                for (int e = lane; e < total4; e += 32) {
                    float4 q = src4[e];
                    chk_raw = fmaf(q.x, 0.0f, fmaf(q.y, 0.0f, fmaf(q.z, 0.0f, fmaf(q.w, 0.0f, chk_raw))));
                    if (norm) {
                        const float nf = nfs[v];
                        q.x = __fmul_rn(q.x, nf); q.y = __fmul_rn(q.y, nf); q.z = __fmul_rn(q.z, nf); q.w = __fmul_rn(q.w, nf);
                        chk_out = fmaf(q.x, 0.0f, fmaf(q.y, 0.0f, fmaf(q.z, 0.0f, fmaf(q.w, 0.0f, chk_out))));
                    }
                    q.x = q.x < 0.0f ? 0.0f : q.x; q.y = q.y < 0.0f ? 0.0f : q.y;  // y[y < 0] = 0 (core.py:452)
                    q.z = q.z < 0.0f ? 0.0f : q.z; q.w = q.w < 0.0f ? 0.0f : q.w;
                    __stcs(out4 + e, q);
                    j4 += 32;
                    while (j4 >= q4) { j4 -= q4; ++v; }
                }
                if (chk_raw != 0.0f) st |= 1u;        // NaN != 0
                else if (chk_out != 0.0f) st |= 2u;
            }
        } else if (nkeep) {
            const long long blk = v0 / PRE_BLOCK_VOX;
            long long pos0 = v0;
            if (p.mask) {  // rank of the chunk's first kept voxel: block prefix + kept voxels of the block before the chunk
                const long long b0v = blk * PRE_BLOCK_VOX;
                int before = 0;
                for (long long v = b0v + lane; v < v0; v += 32) before += (p.mask[v] == 1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(FULLM, before, o);
                pos0 = p.block_off[blk] + before;
            }
            if (pos0 + nkeep > p.y_cap) {
                st |= 4u;
            } else {
                if (keep) p.vox_idx[pos0 + __popc(keepmask & ((1u << lane) - 1u))] = (int)(v0 + lane);
                float *out = p.y + pos0 * m_out;
                const int total = nkeep * m_out;
                int rr = 0, j = lane;
                while (j >= m_out) { j -= m_out; ++rr; }
                bool bad = false;
#pragma unroll 2
                for (int e = lane; e < total; e += 32) {
                    const int v = keepmask == FULLM ? rr : (int)__fns(keepmask, 0, rr + 1);
                    float f;
                    if (diravg) {
                        f = avg[v * m_out + j];
                    } else if (merge && j == 0) {
                        f = mb0s[v];
                    } else {
                        f = raw[v * nS + (merge ? p.dwi_idx[j - 1] : j)];
                        if (norm) f = __fmul_rn(f, nfs[v]);
                    }
                    if (!isfinite(f)) { bad = true; if (replace) f = p.repl; }
                    __stcs(out + e, f < 0.0f ? 0.0f : f);  // y[y < 0] = 0 (core.py:452)
                    j += 32;
                    while (j >= m_out) { j -= m_out; ++rr; }
                }
                if (bad) st |= 2u;
            }
        }
    }
    st = __reduce_or_sync(FULLM, st);
    if (lane == 0 && st) atomicOr(p.status, st);
}

// mean b0 only (amx_mean_b0): same staging, one value per voxel out
__global__ void __launch_bounds__(128) k_mean_b0(const float *__restrict__ dwi, long long n_total, int nS, int stride,
                                                 const int *__restrict__ b0_idx, int b0_count, float *mean_b0s, long long n_chunks)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float *rows = smem + (size_t)warp * (CH * stride + CH);
    for (long long c = (long long)blockIdx.x * wpb + warp; c < n_chunks; c += (long long)gridDim.x * wpb) {
        const long long v0 = c * CH;
        const int nvox = (int)min((long long)CH, n_total - v0);
        stage_rows(rows, dwi + v0 * nS, nvox * nS, nS, stride, lane, false, 0.0f);
        __syncwarp();
        if (lane < nvox) {
            const float *r = rows + lane * stride;
            mean_b0s[v0 + lane] = seq_mean(b0_count, [&](int i) { return r[b0_idx[i]]; });
        }
        __syncwarp();
    }
}

// ---- DTI principal direction ---------------------------------------------------------------------------------------
// cyclic Jacobi on the symmetric 3x3 tensor (fp64): a = {xx, xy, yy, xz, yz, zz}; returns the unit eigenvector of the
// largest eigenvalue (the column dipy's decompose_tensor puts first, dipy/reconst/dti.py `decompose_tensor`)
__device__ __forceinline__ void principal_evec(double a00, double a01, double a11, double a02, double a12, double a22, double *out)
{
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
        const double dg = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-18 * dg) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int pI = pq == 2 ? 1 : 0, qI = pq == 0 ? 1 : 2;
            const double apq = A[pI][qI];
            if (apq == 0.0) continue;
            const double theta = (A[qI][qI] - A[pI][pI]) / (2.0 * apq);
            double t;
            if (fabs(theta) > 1e150) t = 0.5 / theta;
            else t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
            const int rI = 3 - pI - qI;
            const double app = A[pI][pI], aqq = A[qI][qI];
            A[pI][pI] = app - t * apq;
            A[qI][qI] = aqq + t * apq;
            A[pI][qI] = A[qI][pI] = 0.0;
            const double arp = A[rI][pI], arq = A[rI][qI];
            A[rI][pI] = A[pI][rI] = cs * arp - sn * arq;
            A[rI][qI] = A[qI][rI] = sn * arp + cs * arq;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vp = V[k][pI], vq = V[k][qI];
                V[k][pI] = cs * vp - sn * vq;
                V[k][qI] = sn * vp + cs * vq;
            }
        }
    }
    int b = 0;
    if (A[1][1] > A[b][b]) b = 1;
    if (A[2][2] > (b == 0 ? A[0][0] : A[1][1])) b = 2;
    double x = b == 0 ? V[0][0] : b == 1 ? V[0][1] : V[0][2];
    double y = b == 0 ? V[1][0] : b == 1 ? V[1][1] : V[1][2];
    double z = b == 0 ? V[2][0] : b == 1 ? V[2][1] : V[2][2];
    const double nrm = sqrt(x * x + y * y + z * z);
    out[0] = x / nrm; out[1] = y / nrm; out[2] = z / nrm;
}

// W (6 x m, fp64) sits in shared memory; one THREAD per voxel reads its row straight from global memory in 16-byte pieces
// (a warp touches 32 different 128-byte lines per request and consumes each line completely over the next requests, so
// L1 turns the strided pattern into full-line DRAM reads; no staging pass, 32 resident warps per SM for the fp64 log chains).
template <typename T>
__global__ void __launch_bounds__(256) k_dti(const T *__restrict__ y, long long n_vox, int m, const double *__restrict__ W,
                                             double min_signal, double *dirs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sW = reinterpret_cast<double *>(smem_raw);
    for (int i = threadIdx.x; i < 6 * m; i += blockDim.x) sW[i] = W[i];
    __syncthreads();
    constexpr int VEC = 16 / sizeof(T);  // elements per 16-byte load
    const bool vec_ok = (m % VEC) == 0 && ((uintptr_t)y & 15) == 0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vox; v += (long long)gridDim.x * blockDim.x) {
        const T *r = y + v * m;
        double d[6] = {0, 0, 0, 0, 0, 0};
        if (vec_ok) {
#pragma unroll 1
            for (int j = 0; j < m; j += VEC) {
                T x[VEC];
                *reinterpret_cast<int4 *>(x) = __ldg(reinterpret_cast<const int4 *>(r + j));
                double ls[VEC];
#pragma unroll
                for (int t = 0; t < VEC; ++t) ls[t] = log(fmax((double)x[t], min_signal));  // independent chains
#pragma unroll
                for (int t = 0; t < VEC; ++t) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) d[k] = fma(sW[k * m + j + t], ls[t], d[k]);
                }
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < m; ++j) {
                const double ls = log(fmax((double)r[j], min_signal));
#pragma unroll
                for (int k = 0; k < 6; ++k) d[k] = fma(sW[k * m + j], ls, d[k]);
            }
        }
        double e[3];
        principal_evec(d[0], d[1], d[2], d[3], d[4], d[5], e);
        double *o = dirs + v * 3;
        o[0] = e[0]; o[1] = e[1]; o[2] = e[2];
    }
}

// WLS variant (DTI_fit_method 'WLS', amico/core.py:95, 419, 436 -> dipy wls_fit_tensor): weights w = exp(X beta_ols) -- the OLS
// prediction of the signal --, then beta = argmin || diag(w) (X beta - log s) ||, here through the 7 x 7 normal equations
// X^T W^2 X beta = X^T W^2 log s (Cholesky, columns of X scaled to unit size; dipy takes pinv(diag(w) X): same minimiser).
// One thread per voxel, two passes over its row; W7 (7 x m rows of pinv(X)) and X (m x 7) in shared memory.
template <typename T>
__global__ void __launch_bounds__(128) k_dti_wls(const T *__restrict__ y, long long n_vox, int m, const double *__restrict__ W7,
                                                 const double *__restrict__ X, double min_signal, double *dirs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sW = reinterpret_cast<double *>(smem_raw), *sX = sW + 7 * m;
    __shared__ double cs[7];
    for (int i = threadIdx.x; i < 7 * m; i += blockDim.x) { sW[i] = W7[i]; sX[i] = X[i]; }
    __syncthreads();
    if (threadIdx.x < 7) {  // column scale: 1 / max |X[:, k]|
        double mx = 0.0;
        for (int j = 0; j < m; ++j) mx = fmax(mx, fabs(sX[j * 7 + threadIdx.x]));
        cs[threadIdx.x] = mx > 0.0 ? 1.0 / mx : 1.0;
    }
    __syncthreads();
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vox; v += (long long)gridDim.x * blockDim.x) {
        const T *r = y + v * m;
        double b[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
        for (int j = 0; j < m; ++j) {
            const double ls = log(fmax((double)r[j], min_signal));
#pragma unroll
            for (int k = 0; k < 7; ++k) b[k] = fma(sW[k * m + j], ls, b[k]);
        }
        double N[28], rhs[7];  // lower triangle of X^T W^2 X (scaled), row-major packed
#pragma unroll
        for (int k = 0; k < 28; ++k) N[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) rhs[k] = 0.0;
#pragma unroll 1
        for (int j = 0; j < m; ++j) {
            const double ls = log(fmax((double)r[j], min_signal));
            double xs[7], yh = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) { const double x = sX[j * 7 + k]; yh = fma(x, b[k], yh); xs[k] = x * cs[k]; }
            const double w = exp(yh), w2 = w * w;
#pragma unroll
            for (int a = 0; a < 7; ++a) {
                const double wa = w2 * xs[a];
                rhs[a] = fma(wa, ls, rhs[a]);
#pragma unroll
                for (int c = 0; c <= a; ++c) N[a * (a + 1) / 2 + c] = fma(wa, xs[c], N[a * (a + 1) / 2 + c]);
            }
        }
        // Cholesky N = L L^T in place, then two triangular solves
#pragma unroll
        for (int a = 0; a < 7; ++a) {
#pragma unroll
            for (int c = 0; c <= a; ++c) {
                double sum = N[a * (a + 1) / 2 + c];
#pragma unroll
                for (int k = 0; k < c; ++k) sum = fma(-N[a * (a + 1) / 2 + k], N[c * (c + 1) / 2 + k], sum);
                N[a * (a + 1) / 2 + c] = (a == c) ? sqrt(fmax(sum, 1e-300)) : sum / N[c * (c + 1) / 2 + c];
            }
        }
#pragma unroll
        for (int a = 0; a < 7; ++a) {
            double sum = rhs[a];
#pragma unroll
            for (int k = 0; k < a; ++k) sum = fma(-N[a * (a + 1) / 2 + k], rhs[k], sum);
            rhs[a] = sum / N[a * (a + 1) / 2 + a];
        }
#pragma unroll
        for (int a = 6; a >= 0; --a) {
            double sum = rhs[a];
#pragma unroll
            for (int k = a + 1; k < 7; ++k) sum = fma(-N[k * (k + 1) / 2 + a], rhs[k], sum);
            rhs[a] = sum / N[a * (a + 1) / 2 + a];
        }
        double e[3];
        principal_evec(rhs[0] * cs[0], rhs[1] * cs[1], rhs[2] * cs[2], rhs[3] * cs[3], rhs[4] * cs[4], rhs[5] * cs[5], e);
        double *o = dirs + v * 3;
        o[0] = e[0]; o[1] = e[1]; o[2] = e[2];
    }
}

// ---- result scatter ----------------------------------------------------------------------------------------------------
__global__ void k_scatter_maps(const double *__restrict__ values, long long n_vox, int k, const int *__restrict__ vox_idx,
                               float *volume)
{
    const long long total = n_vox * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / k;
        const int c = (int)(e - i * k);
        volume[(long long)vox_idx[i] * k + c] = (float)values[e];
    }
}


// ---- kernel resampling (amico/lut.pyx:274-311) --------------------------------------------------------------------
// out[row][j] = 1 when scheme row merge_idx[j] is not a dwi row, else sum_k Ylm[p][k] * KRlm[row][k] with p the dwi position of
// that row; rows = (atom, direction) pairs.  CTA tile: RT rows x JT output columns; the K tile and the (transposed)
// Ylm tile sit in shared memory; accumulation in fp64, one rounding to fp32 (the reference's np.dot is a float32 BLAS
// gemv whose summation order depends on the BLAS build: this result is within 0.5 ulp of the exact dot product).
constexpr int RS_RT = 32, RS_JT = 32;
__global__ void __launch_bounds__(256) k_resample(const float *__restrict__ KRlm, long long n_rows, int n_coef,
                                                  const float *__restrict__ Ylm, const int *__restrict__ colsrc, int nS_out, float *out)
{
    extern __shared__ __align__(16) float rs_smem[];
    float *Ks = rs_smem;                       // [RS_RT][n_coef]
    float *Ys = Ks + (size_t)RS_RT * n_coef;   // [n_coef][RS_JT]  (transposed: threads of a warp read consecutive words)
    __shared__ int src[RS_JT];
    const long long r0 = (long long)blockIdx.x * RS_RT;
    const int j0 = blockIdx.y * RS_JT;
    const int nr = (int)min((long long)RS_RT, n_rows - r0), nj = min(RS_JT, nS_out - j0);
    if (threadIdx.x < RS_JT) src[threadIdx.x] = threadIdx.x < nj ? colsrc[j0 + threadIdx.x] : -1;
    __syncthreads();
    for (int e = threadIdx.x; e < nr * n_coef; e += blockDim.x) Ks[e] = KRlm[r0 * n_coef + e];
    for (int e = threadIdx.x; e < RS_JT * n_coef; e += blockDim.x) {
        const int jj = e / n_coef, k = e - jj * n_coef;
        const int p = src[jj];
        Ys[k * RS_JT + jj] = p >= 0 ? Ylm[(size_t)p * n_coef + k] : 0.0f;
    }
    __syncthreads();
    const int jj = threadIdx.x & (RS_JT - 1);
    for (int rr = threadIdx.x / RS_JT; rr < nr; rr += blockDim.x / RS_JT) {
        if (jj >= nj) continue;
        float v = 1.0f;  // KR = np.ones(...) (lut.pyx:297): rows outside idx_out (the b0 volumes) stay 1
        if (src[jj] >= 0) {
            const float *kr = Ks + (size_t)rr * n_coef;
            double acc = 0.0;
#pragma unroll 4
            for (int k = 0; k < n_coef; ++k) acc = fma((double)Ys[k * RS_JT + jj], (double)kr[k], acc);
            v = (float)acc;
        }
        out[(r0 + rr) * nS_out + j0 + jj] = v;
    }
}


// ---- on-disk volume -> voxel-major float32 ------------------------------------------------------------------------
// A NIfTI image stores x fastest and the volume index slowest: memory [nS][n_total].  The fit wants each voxel's nS
// values contiguous: [n_total][nS].  Classic tiled transpose through shared memory (both sides coalesced), fused with the
// dtype conversion and the header's slope / intercept (nibabel's get_fdata computes raw * slope + inter in float64; the
// reference then casts to float32, amico/core.py:136): same two roundings here.
template <typename T>
__global__ void __launch_bounds__(256) k_to_voxel_major(const T *__restrict__ src, long long n_total, int nS, double slope, double inter,
                                                        int scale, float *__restrict__ dst)
{
    __shared__ float tile[32][33];
    const long long v0 = (long long)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int sI = s0 + r;
        const long long v = v0 + tx;
        if (sI < nS && v < n_total) {
            const T raw = src[(long long)sI * n_total + v];
            tile[r][tx] = scale ? (float)((double)raw * slope + inter) : (float)raw;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long v = v0 + r;
        const int sI = s0 + tx;
        if (sI < nS && v < n_total) dst[v * nS + sI] = tile[tx][r];
    }
}

// ---- host helpers ------------------------------------------------------------------------------------------------------
struct Tmp {  // stream-ordered temporaries
    cudaStream_t s;
    std::vector<void *> ptrs;
    explicit Tmp(cudaStream_t s_) : s(s_) {}
    ~Tmp() { for (void *p : ptrs) cudaFreeAsync(p, s); }
    cudaError_t alloc(void **p, size_t bytes)
    {
        cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, s);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

int pick_device(int device, int *sm_count, int *max_smem)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return amx::set_error(AMX_E_CUDA, "no CUDA device available (%s): amico_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return amx::set_error(AMX_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    AMX_CK(cudaSetDevice(device));
    {   // stream-ordered temporaries: keep freed blocks in the pool across synchronisations (default: trimmed at every sync)
        static bool pool_set[64];
        if (!pool_set[device & 63]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long keep = 1ull << 30;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_set[device & 63] = true;
        }
    }
    AMX_CK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
    AMX_CK(cudaDeviceGetAttribute(max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    return AMX_OK;
}

// warps per CTA such that `fixed + warps * per_warp` bytes fit the opt-in shared-memory limit (<= 4)
int warps_for(size_t fixed, size_t per_warp, int max_smem)
{
    int w = 4;
    while (w > 0 && fixed + (size_t)w * per_warp > (size_t)max_smem) --w;
    return w;
}

int check_idx(const int32_t *idx, int n, int nS, const char *what)
{
    if (n < 0 || (n > 0 && !idx)) return amx::set_error(AMX_E_INVALID, "%s: bad index list", what);
    for (int i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= nS) return amx::set_error(AMX_E_INVALID, "%s[%d]=%d outside [0,%d)", what, i, idx[i], nS);
    return AMX_OK;
}

}  // namespace

extern "C" {

int amx_preprocess(const amx_pre_args *a, int64_t *n_kept, int *m_out_p)
{
    if (!a || !n_kept) return amx::set_error(AMX_E_INVALID, "NULL argument");
    *n_kept = 0;
    if (!a->dwi || a->n_total <= 0 || a->nS <= 0 || !a->y || !a->vox_idx || a->y_capacity < 0)
        return amx::set_error(AMX_E_INVALID, "bad volume / output arguments");
    if (a->n_total > 0x7fffffffLL) return amx::set_error(AMX_E_INVALID, "more than 2^31 voxels");
    if ((a->flags & AMX_PRE_MERGE_B0) && (a->flags & AMX_PRE_DIR_AVG))
        return amx::set_error(AMX_E_INVALID, "doMergeB0 together with doDirectionalAverage is not meaningful: the reference "
                                             "indexes the merged volume with the un-merged scheme (amico/core.py:225-245)");
    int rc;
    if ((rc = check_idx(a->b0_idx, a->b0_count, a->nS, "b0_idx")) || (rc = check_idx(a->dwi_idx, a->dwi_count, a->nS, "dwi_idx"))) return rc;
    const bool needs_b0 = a->flags & (AMX_PRE_NORMALIZE | AMX_PRE_MERGE_B0 | AMX_PRE_DIR_AVG);
    if (needs_b0 && a->b0_count <= 0)
        return amx::set_error(AMX_E_INVALID, "No b0 volume to normalize signal with");  // core.py:214
    int m_out = a->nS;
    int n_shell_idx = 0;
    if (a->flags & AMX_PRE_MERGE_B0) {
        m_out = 1 + a->dwi_count;
        for (int k = 1; k < a->dwi_count; ++k)
            if (a->dwi_idx[k] <= a->dwi_idx[k - 1]) return amx::set_error(AMX_E_INVALID, "dwi_idx must be ascending");
    }
    if (a->flags & AMX_PRE_DIR_AVG) {
        if (a->n_shells <= 0 || !a->shell_off || !a->shell_idx || a->n_shells + 1 > a->nS)
            return amx::set_error(AMX_E_INVALID, "bad shell lists");
        n_shell_idx = a->shell_off[a->n_shells];
        for (int s = 0; s < a->n_shells; ++s)
            if (a->shell_off[s + 1] <= a->shell_off[s] || a->shell_off[0] != 0) return amx::set_error(AMX_E_INVALID, "bad shell offsets");
        if ((rc = check_idx(a->shell_idx, n_shell_idx, a->nS, "shell_idx"))) return rc;
        m_out = 1 + a->n_shells;
    }
    if (m_out_p) *m_out_p = m_out;
    int sm = 0, max_smem = 0;
    if ((rc = pick_device(a->device, &sm, &max_smem))) return rc;
    cudaStream_t s = (cudaStream_t)a->stream;
    Tmp tmp(s);
    const bool host = a->space == AMX_SPACE_HOST;
    if (!host && a->space != AMX_SPACE_DEVICE) return amx::set_error(AMX_E_INVALID, "bad space");

    PreParams p{};
    p.n_total = a->n_total; p.nS = a->nS;
    p.b0_count = a->b0_count; p.dwi_count = a->dwi_count; p.n_shells = (a->flags & AMX_PRE_DIR_AVG) ? a->n_shells : 0;
    p.flags = a->flags; p.thr = a->b0_threshold; p.repl = a->replace_bad; p.m_out = m_out; p.y_cap = a->y_capacity;

    // small index lists -> one device buffer
    std::vector<int> idx;
    idx.insert(idx.end(), a->b0_idx, a->b0_idx + a->b0_count);
    const size_t o_dwi = idx.size();
    idx.insert(idx.end(), a->dwi_idx, a->dwi_idx + a->dwi_count);
    const size_t o_sh = idx.size();
    if (p.n_shells) idx.insert(idx.end(), a->shell_idx, a->shell_idx + n_shell_idx);
    const size_t o_off = idx.size();
    if (p.n_shells) idx.insert(idx.end(), a->shell_off, a->shell_off + p.n_shells + 1);
    int *d_idx = nullptr;
    AMX_CK(tmp.alloc((void **)&d_idx, idx.size() * sizeof(int)));
    if (!idx.empty()) AMX_CK(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    p.b0_idx = d_idx; p.dwi_idx = d_idx + o_dwi; p.shell_idx = d_idx + o_sh; p.shell_off = d_idx + o_off;

    const size_t vol_bytes = (size_t)a->n_total * a->nS * sizeof(float);
    float *d_dwi = nullptr, *d_y = nullptr, *d_mb = nullptr;
    unsigned char *d_mask = nullptr;
    int *d_vidx = nullptr;
    if (host) {
        AMX_CK(tmp.alloc((void **)&d_dwi, vol_bytes));
        AMX_CK(cudaMemcpyAsync(d_dwi, a->dwi, vol_bytes, cudaMemcpyHostToDevice, s));
        if (a->mask) {
            AMX_CK(tmp.alloc((void **)&d_mask, (size_t)a->n_total));
            AMX_CK(cudaMemcpyAsync(d_mask, a->mask, (size_t)a->n_total, cudaMemcpyHostToDevice, s));
        }
        AMX_CK(tmp.alloc((void **)&d_y, (size_t)a->y_capacity * m_out * sizeof(float)));
        AMX_CK(tmp.alloc((void **)&d_vidx, (size_t)a->y_capacity * sizeof(int)));
        if (a->mean_b0s) AMX_CK(tmp.alloc((void **)&d_mb, (size_t)a->n_total * sizeof(float)));
        p.dwi = d_dwi; p.mask = d_mask; p.y = d_y; p.vox_idx = d_vidx; p.mean_b0s = d_mb;
    } else {
        p.dwi = a->dwi; p.mask = a->mask; p.y = a->y; p.vox_idx = a->vox_idx; p.mean_b0s = a->mean_b0s;
    }
    if (((uintptr_t)p.dwi & 15) != 0) return amx::set_error(AMX_E_INVALID, "dwi must be 16-byte aligned");

    const int n_blocks = (int)((a->n_total + PRE_BLOCK_VOX - 1) / PRE_BLOCK_VOX);
    long long *d_counts = nullptr;
    unsigned *d_status = nullptr;
    AMX_CK(tmp.alloc((void **)&d_counts, (size_t)(n_blocks + 1) * sizeof(long long)));
    AMX_CK(tmp.alloc((void **)&d_status, 16));
    AMX_CK(cudaMemsetAsync(d_status, 0, 16, s));
    if (p.mask) {  // order-preserving compaction: kept voxels per 1024-voxel block, exclusive scan
        k_mask_count<<<n_blocks, 256, 0, s>>>(p.mask, a->n_total, d_counts);
        k_scan_counts<<<1, 1024, 0, s>>>(d_counts, n_blocks);
    }
    p.block_off = p.mask ? d_counts : nullptr;  // no mask: block b starts at row 1024 b
    p.status = d_status;

    const size_t per_warp = pre_warp_bytes(a->nS, m_out, a->flags & AMX_PRE_DIR_AVG);
    const int warps = warps_for(1024, per_warp, max_smem);
    if (warps <= 0) return amx::set_error(AMX_E_INVALID, "nS=%d too large for the shared-memory staging", a->nS);
    const size_t smem = (size_t)warps * per_warp;
    AMX_CK(cudaFuncSetAttribute(k_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_chunks = (a->n_total + CH - 1) / CH;
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)max_smem / (smem + 1024)));
    const long long want = (n_chunks + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm * ctas_per_sm));
    k_preprocess<<<grid, warps * 32, smem, s>>>(p, n_chunks, (unsigned)per_warp);
    AMX_CK(cudaGetLastError());

    long long total = 0;
    unsigned status = 0;
    if (p.mask) AMX_CK(cudaMemcpyAsync(&total, d_counts + n_blocks, sizeof total, cudaMemcpyDeviceToHost, s));
    else total = a->n_total;
    AMX_CK(cudaMemcpyAsync(&status, d_status, sizeof status, cudaMemcpyDeviceToHost, s));
    AMX_CK(cudaStreamSynchronize(s));
    *n_kept = total;
    if (status & 4u) return amx::set_error(AMX_E_INVALID, "y_capacity=%lld is smaller than the %lld mask voxels", (long long)a->y_capacity, total);
    if ((status & 1u) && !(a->flags & AMX_PRE_REPLACE_BAD))
        return amx::set_error(AMX_E_NONFINITE, "Nan or Inf values in the raw signal. Try using the \"replace_bad_voxels\" or "
                                               "\"b0_min_signal\" parameters when calling \"load_data()\"");
    if ((status & 2u) && !(a->flags & AMX_PRE_REPLACE_BAD))
        return amx::set_error(AMX_E_NONFINITE, "Nan or Inf values in the signal after the pre-processing. Try using the "
                                               "\"replace_bad_voxels\" or \"b0_min_signal\" parameters when calling \"load_data()\"");
    if (host) {
        AMX_CK(cudaMemcpyAsync(a->y, d_y, (size_t)total * m_out * sizeof(float), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaMemcpyAsync(a->vox_idx, d_vidx, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (a->mean_b0s) AMX_CK(cudaMemcpyAsync(a->mean_b0s, d_mb, (size_t)a->n_total * sizeof(float), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaStreamSynchronize(s));
    }
    return AMX_OK;
}

int amx_mean_b0(int space, int device, const float *dwi, int64_t n_total, int nS, const int32_t *b0_idx, int b0_count,
                float *mean_b0s, void *stream)
{
    if (!dwi || !mean_b0s || n_total <= 0 || nS <= 0) return amx::set_error(AMX_E_INVALID, "bad arguments");
    if (b0_count <= 0) return amx::set_error(AMX_E_INVALID, "No b0 volume to normalize signal with");
    int rc;
    if ((rc = check_idx(b0_idx, b0_count, nS, "b0_idx"))) return rc;
    int sm = 0, max_smem = 0;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    int *d_idx = nullptr;
    AMX_CK(tmp.alloc((void **)&d_idx, (size_t)b0_count * sizeof(int)));
    AMX_CK(cudaMemcpyAsync(d_idx, b0_idx, (size_t)b0_count * sizeof(int), cudaMemcpyHostToDevice, s));
    const float *src = dwi;
    float *dst = mean_b0s;
    if (host) {
        float *d_dwi = nullptr, *d_mb = nullptr;
        AMX_CK(tmp.alloc((void **)&d_dwi, (size_t)n_total * nS * sizeof(float)));
        AMX_CK(cudaMemcpyAsync(d_dwi, dwi, (size_t)n_total * nS * sizeof(float), cudaMemcpyHostToDevice, s));
        AMX_CK(tmp.alloc((void **)&d_mb, (size_t)n_total * sizeof(float)));
        src = d_dwi; dst = d_mb;
    }
    const int stride = nS | 1;
    const size_t per_warp = (size_t)(CH * stride + CH) * sizeof(float);
    const int warps = warps_for(0, per_warp, max_smem);
    if (warps <= 0) return amx::set_error(AMX_E_INVALID, "nS=%d too large for the shared-memory staging", nS);
    const size_t smem = (size_t)warps * per_warp;
    AMX_CK(cudaFuncSetAttribute(k_mean_b0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_chunks = (n_total + CH - 1) / CH;
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)max_smem / smem));
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_chunks + warps - 1) / warps, (long long)sm * ctas_per_sm));
    k_mean_b0<<<grid, warps * 32, smem, s>>>(src, n_total, nS, stride, d_idx, b0_count, dst, n_chunks);
    AMX_CK(cudaGetLastError());
    if (host) AMX_CK(cudaMemcpyAsync(mean_b0s, dst, (size_t)n_total * sizeof(float), cudaMemcpyDeviceToHost, s));
    AMX_CK(cudaStreamSynchronize(s));
    return AMX_OK;
}

int amx_dti_directions(int device, int space, const void *y, int y_dtype, int64_t n_vox, int m, const double *W,
                       double min_signal, double *dirs, void *stream)
{
    if (!y || !W || !dirs || n_vox < 0 || m <= 0) return amx::set_error(AMX_E_INVALID, "bad arguments");
    if (y_dtype != AMX_F32 && y_dtype != AMX_F64) return amx::set_error(AMX_E_INVALID, "bad y_dtype");
    int rc, sm = 0, max_smem = 0;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    if (n_vox == 0) return AMX_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    const size_t esz = y_dtype == AMX_F64 ? 8 : 4;
    double *d_W = nullptr;
    AMX_CK(tmp.alloc((void **)&d_W, (size_t)6 * m * sizeof(double)));
    AMX_CK(cudaMemcpyAsync(d_W, W, (size_t)6 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    const void *src = y;
    double *dst = dirs;
    if (host) {
        void *d_y = nullptr;
        double *d_d = nullptr;
        AMX_CK(tmp.alloc(&d_y, (size_t)n_vox * m * esz));
        AMX_CK(cudaMemcpyAsync(d_y, y, (size_t)n_vox * m * esz, cudaMemcpyHostToDevice, s));
        AMX_CK(tmp.alloc((void **)&d_d, (size_t)n_vox * 3 * sizeof(double)));
        src = d_y; dst = d_d;
    }
    const size_t smem = (size_t)6 * m * sizeof(double);
    if (smem > (size_t)max_smem) return amx::set_error(AMX_E_INVALID, "m=%d too large for the shared-memory weight table", m);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_vox + 255) / 256, (long long)sm * 8));
    if (y_dtype == AMX_F64) {
        AMX_CK(cudaFuncSetAttribute(k_dti<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dti<double><<<grid, 256, smem, s>>>((const double *)src, n_vox, m, d_W, min_signal, dst);
    } else {
        AMX_CK(cudaFuncSetAttribute(k_dti<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dti<float><<<grid, 256, smem, s>>>((const float *)src, n_vox, m, d_W, min_signal, dst);
    }
    AMX_CK(cudaGetLastError());
    if (host) {
        AMX_CK(cudaMemcpyAsync(dirs, dst, (size_t)n_vox * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaStreamSynchronize(s));
    }
    return AMX_OK;
}

int amx_dti_directions_wls(int device, int space, const void *y, int y_dtype, int64_t n_vox, int m, const double *W7, const double *X,
                           double min_signal, double *dirs, void *stream)
{
    if (!y || !W7 || !X || !dirs || n_vox < 0 || m <= 0) return amx::set_error(AMX_E_INVALID, "bad arguments");
    if (y_dtype != AMX_F32 && y_dtype != AMX_F64) return amx::set_error(AMX_E_INVALID, "bad y_dtype");
    int rc, sm = 0, max_smem = 0;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    if (n_vox == 0) return AMX_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    const size_t esz = y_dtype == AMX_F64 ? 8 : 4;
    double *d_W = nullptr, *d_X = nullptr;
    AMX_CK(tmp.alloc((void **)&d_W, (size_t)7 * m * sizeof(double)));
    AMX_CK(tmp.alloc((void **)&d_X, (size_t)7 * m * sizeof(double)));
    AMX_CK(cudaMemcpyAsync(d_W, W7, (size_t)7 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    AMX_CK(cudaMemcpyAsync(d_X, X, (size_t)7 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    const void *src = y;
    double *dst = dirs;
    if (host) {
        void *d_y = nullptr;
        double *d_d = nullptr;
        AMX_CK(tmp.alloc(&d_y, (size_t)n_vox * m * esz));
        AMX_CK(cudaMemcpyAsync(d_y, y, (size_t)n_vox * m * esz, cudaMemcpyHostToDevice, s));
        AMX_CK(tmp.alloc((void **)&d_d, (size_t)n_vox * 3 * sizeof(double)));
        src = d_y; dst = d_d;
    }
    const size_t smem = (size_t)14 * m * sizeof(double);
    if (smem > (size_t)max_smem) return amx::set_error(AMX_E_INVALID, "m=%d too large for the shared-memory design matrix", m);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_vox + 127) / 128, (long long)sm * 8));
    if (y_dtype == AMX_F64) {
        AMX_CK(cudaFuncSetAttribute(k_dti_wls<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dti_wls<double><<<grid, 128, smem, s>>>((const double *)src, n_vox, m, d_W, d_X, min_signal, dst);
    } else {
        AMX_CK(cudaFuncSetAttribute(k_dti_wls<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_dti_wls<float><<<grid, 128, smem, s>>>((const float *)src, n_vox, m, d_W, d_X, min_signal, dst);
    }
    AMX_CK(cudaGetLastError());
    if (host) {
        AMX_CK(cudaMemcpyAsync(dirs, dst, (size_t)n_vox * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaStreamSynchronize(s));
    } else {
        AMX_CK(cudaStreamSynchronize(s));  // W7 / X are host memory read by asynchronous copies
    }
    return AMX_OK;
}

int amx_scatter_maps(int device, int space, const double *values, int64_t n_vox, int k, const int32_t *vox_idx, float *volume,
                     int64_t n_total, void *stream)
{
    if (!volume || n_total <= 0 || k <= 0 || n_vox < 0 || (n_vox > 0 && (!values || !vox_idx)))
        return amx::set_error(AMX_E_INVALID, "bad arguments");
    int rc, sm = 0, max_smem = 0;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    const double *src = values;
    const int *idx = vox_idx;
    float *dst = volume;
    if (host) {
        double *d_v = nullptr;
        int *d_i = nullptr;
        float *d_o = nullptr;
        AMX_CK(tmp.alloc((void **)&d_v, (size_t)n_vox * k * sizeof(double)));
        AMX_CK(tmp.alloc((void **)&d_i, (size_t)n_vox * sizeof(int)));
        AMX_CK(tmp.alloc((void **)&d_o, (size_t)n_total * k * sizeof(float)));
        if (n_vox) {
            AMX_CK(cudaMemcpyAsync(d_v, values, (size_t)n_vox * k * sizeof(double), cudaMemcpyHostToDevice, s));
            AMX_CK(cudaMemcpyAsync(d_i, vox_idx, (size_t)n_vox * sizeof(int), cudaMemcpyHostToDevice, s));
        }
        src = d_v; idx = d_i; dst = d_o;
    }
    AMX_CK(cudaMemsetAsync(dst, 0, (size_t)n_total * k * sizeof(float), s));
    if (n_vox) {
        const long long total = (long long)n_vox * k;
        const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm * 8));
        k_scatter_maps<<<grid, 256, 0, s>>>(src, n_vox, k, idx, dst);
        AMX_CK(cudaGetLastError());
    }
    if (host) {
        AMX_CK(cudaMemcpyAsync(volume, dst, (size_t)n_total * k * sizeof(float), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaStreamSynchronize(s));
    }
    return AMX_OK;
}

int amx_resample_kernels(int device, int space, const float *KRlm, int64_t n_rows, int n_coef, const float *Ylm_out,
                         const int32_t *idx_out, int dwi_count, const int32_t *merge_idx, int nS_out, int nS, float *out,
                         void *stream)
{
    if (!KRlm || !Ylm_out || !idx_out || !merge_idx || !out || n_rows <= 0 || n_coef <= 0 || dwi_count <= 0 || nS_out <= 0 || nS <= 0)
        return amx::set_error(AMX_E_INVALID, "bad arguments");
    int rc, sm = 0, max_smem = 0;
    if ((rc = check_idx(idx_out, dwi_count, nS, "idx_out")) || (rc = check_idx(merge_idx, nS_out, nS, "merge_idx"))) return rc;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    const size_t smem = (size_t)(RS_RT + RS_JT) * n_coef * sizeof(float);
    if (smem > (size_t)max_smem) return amx::set_error(AMX_E_INVALID, "n_coef=%d too large for the shared-memory tiles", n_coef);
    // output column j <- dwi position of scheme row merge_idx[j] (or -1: stays 1)
    std::vector<int> pos(nS, -1), colsrc(nS_out);
    for (int p = 0; p < dwi_count; ++p) pos[idx_out[p]] = p;
    for (int j = 0; j < nS_out; ++j) colsrc[j] = pos[merge_idx[j]];
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    int *d_col = nullptr;
    AMX_CK(tmp.alloc((void **)&d_col, (size_t)nS_out * sizeof(int)));
    AMX_CK(cudaMemcpyAsync(d_col, colsrc.data(), (size_t)nS_out * sizeof(int), cudaMemcpyHostToDevice, s));
    const float *d_K = KRlm, *d_Y = Ylm_out;
    float *d_o = out;
    if (host) {
        float *k = nullptr, *yl = nullptr, *o = nullptr;
        AMX_CK(tmp.alloc((void **)&k, (size_t)n_rows * n_coef * sizeof(float)));
        AMX_CK(tmp.alloc((void **)&yl, (size_t)dwi_count * n_coef * sizeof(float)));
        AMX_CK(tmp.alloc((void **)&o, (size_t)n_rows * nS_out * sizeof(float)));
        AMX_CK(cudaMemcpyAsync(k, KRlm, (size_t)n_rows * n_coef * sizeof(float), cudaMemcpyHostToDevice, s));
        AMX_CK(cudaMemcpyAsync(yl, Ylm_out, (size_t)dwi_count * n_coef * sizeof(float), cudaMemcpyHostToDevice, s));
        d_K = k; d_Y = yl; d_o = o;
    }
    AMX_CK(cudaFuncSetAttribute(k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n_rows + RS_RT - 1) / RS_RT), (unsigned)((nS_out + RS_JT - 1) / RS_JT));
    k_resample<<<grid, 256, smem, s>>>(d_K, n_rows, n_coef, d_Y, d_col, nS_out, d_o);
    AMX_CK(cudaGetLastError());
    if (host) AMX_CK(cudaMemcpyAsync(out, d_o, (size_t)n_rows * nS_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    AMX_CK(cudaStreamSynchronize(s));
    return AMX_OK;
}

int amx_volume_to_voxel_major(int device, int space, const void *src, int nifti_datatype, int64_t n_total, int nS, double scl_slope,
                              double scl_inter, float *dst, void *stream)
{
    if (!src || !dst || n_total <= 0 || nS <= 0) return amx::set_error(AMX_E_INVALID, "bad arguments");
    size_t esz;
    switch (nifti_datatype) {
        case 2: case 256: esz = 1; break;
        case 4: case 512: esz = 2; break;
        case 8: case 16: case 768: esz = 4; break;
        case 64: esz = 8; break;
        default: return amx::set_error(AMX_E_INVALID, "unsupported NIfTI datatype %d", nifti_datatype);
    }
    int rc, sm = 0, max_smem = 0;
    if ((rc = pick_device(device, &sm, &max_smem))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Tmp tmp(s);
    const bool host = space == AMX_SPACE_HOST;
    const void *d_src = src;
    float *d_dst = dst;
    if (host) {
        void *a = nullptr;
        float *o = nullptr;
        AMX_CK(tmp.alloc(&a, (size_t)n_total * nS * esz));
        AMX_CK(tmp.alloc((void **)&o, (size_t)n_total * nS * sizeof(float)));
        AMX_CK(cudaMemcpyAsync(a, src, (size_t)n_total * nS * esz, cudaMemcpyHostToDevice, s));
        d_src = a; d_dst = o;
    }
    // nibabel applies the scaling only when it is meaningful (finite, slope != 0, not the identity)
    const int scale = std::isfinite(scl_slope) && std::isfinite(scl_inter) && scl_slope != 0.0 && (scl_slope != 1.0 || scl_inter != 0.0);
    dim3 grid((unsigned)((n_total + 31) / 32), (unsigned)((nS + 31) / 32));
#define AMX_TVM(T) k_to_voxel_major<T><<<grid, 256, 0, s>>>((const T *)d_src, n_total, nS, scl_slope, scl_inter, scale, d_dst)
    switch (nifti_datatype) {
        case 2: AMX_TVM(unsigned char); break;
        case 256: AMX_TVM(signed char); break;
        case 4: AMX_TVM(short); break;
        case 512: AMX_TVM(unsigned short); break;
        case 8: AMX_TVM(int); break;
        case 768: AMX_TVM(unsigned int); break;
        case 16: AMX_TVM(float); break;
        default: AMX_TVM(double); break;
    }
#undef AMX_TVM
    AMX_CK(cudaGetLastError());
    if (host) {
        AMX_CK(cudaMemcpyAsync(dst, d_dst, (size_t)n_total * nS * sizeof(float), cudaMemcpyDeviceToHost, s));
        AMX_CK(cudaStreamSynchronize(s));
    }
    return AMX_OK;
}

}  // extern "C"
