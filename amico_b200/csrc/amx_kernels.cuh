// Device kernels of the per-voxel fit: LUT index + binning, table construction, fused fit.
#pragma once
#include "amx_solvers.cuh"
#ifndef AMX_GEMM_INLINE
#define AMX_GEMM_INLINE __noinline__
#endif

namespace amx {

enum { MODEL_NODDI = 0, MODEL_FREEWATER = 1, MODEL_CZB = 2, MODEL_SANDI = 3 };
enum { FLAG_RMSE = 1, FLAG_NRMSE = 2, FLAG_EXTRA = 4, FLAG_EXACT = 8 };

// ------------------------------------------------------------------------------------------------
// direction -> LUT index (amico/lut.pyx:314-356), flips `d` in place; -1 when out of range.
__device__ __forceinline__ int dir_to_lut_idx(double *d, const int16_t *__restrict__ htable)
{
    const double PI = 3.14159265358979323846;
    double x = d[0], y = d[1], z = d[2];
    if (y < 0.0) {
        x = -x; y = -y; z = -z;
        d[0] = x; d[1] = y; d[2] = z;
    }
    double i1, i2 = fmod(atan2(y, x), 2.0 * PI);
    if (i2 < 0.0) i2 = fmod(i2 + 2.0 * PI, 2.0 * PI);
    if (i2 > PI) {
        i2 = fmod(atan2(-y, -x), 2.0 * PI);
        i1 = atan2(sqrt(x * x + y * y), -z);
    } else {
        i1 = atan2(sqrt(x * x + y * y), z);
    }
    double r1 = round(i1 / PI * 180.0), r2 = round(i2 / PI * 180.0);
    if (!(r1 >= 0.0 && r1 <= 180.0 && r2 >= 0.0 && r2 <= 180.0)) return -1;
    return (int)htable[(int)r1 * 181 + (int)r2];
}

// status words: [0] error flag, [1] first offending voxel, [2] voxels queued for the scalar slow path,
// [3] voxels a kernel WITHOUT slow path could not finish (-> AMX_E_CAPACITY), [4] voxels queued for the exact path
__global__ void k_lut(double *dirs, long long n, const int16_t *__restrict__ htable, int ndirs, int *lut, int *hist,
                      long long *status, long long vox_offset)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int idx = dir_to_lut_idx(dirs + 3 * i, htable);
    if (idx < 0 || idx >= ndirs) {
        idx = -1;
        atomicExch((unsigned long long *)&status[0], 1ull);
        atomicMin(&status[1], i + vox_offset);
    } else if (hist) {
        atomicAdd(&hist[idx], 1);
    }
    lut[i] = idx;
}

// exclusive scan of the per-direction histogram + tile bookkeeping; one block.
// offs[d] = first slot of bin d in `order`; tile_offs[d] = first tile of bin d; totals[0] = n tiles
__global__ void k_scan_bins(const int *hist, int ndirs, int tile_v, int *offs, int *cursor, int *tile_offs, int *totals)
{
    __shared__ int s_part[1024], s_tpart[1024];
    int tid = threadIdx.x, nt = blockDim.x;
    int per = (ndirs + nt - 1) / nt;
    int b0 = tid * per, b1 = min(ndirs, b0 + per);
    int sum = 0, tsum = 0;
    #pragma unroll 1
    for (int d = b0; d < b1; ++d) { sum += hist[d]; tsum += (hist[d] + tile_v - 1) / tile_v; }
    s_part[tid] = sum; s_tpart[tid] = tsum;
    __syncthreads();
    if (tid == 0) {
        int a = 0, t = 0;
        #pragma unroll 1
        for (int i = 0; i < nt; ++i) {
            int v = s_part[i], tv = s_tpart[i];
            s_part[i] = a; s_tpart[i] = t;
            a += v; t += tv;
        }
        totals[0] = t;
        totals[1] = a;
    }
    __syncthreads();
    int a = s_part[tid], t = s_tpart[tid];
    #pragma unroll 1
    for (int d = b0; d < b1; ++d) {
        offs[d] = a; cursor[d] = a; tile_offs[d] = t;
        a += hist[d]; t += (hist[d] + tile_v - 1) / tile_v;
    }
}

__global__ void k_scatter(const int *lut, long long n, int *cursor, int *order)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = lut[i];
    if (d < 0) return;
    int pos = atomicAdd(&cursor[d], 1);
    order[pos] = (int)i;
}

// balanced: the ceil(c / tile_v) tiles of a bin get (nearly) equal sizes instead of one short tail tile
__global__ void k_tiles(const int *hist, const int *offs, const int *tile_offs, int ndirs, int tile_v, int4 *tiles, int balanced)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndirs) return;
    int c = hist[d], o = offs[d], t = tile_offs[d];
    if (balanced && c > 0) {
        const int nt = (c + tile_v - 1) / tile_v;
        tile_v = (c + nt - 1) / nt;
    }
    #pragma unroll 1
    for (int s = 0; s < c; s += tile_v, ++t) tiles[t] = make_int4(d, o + s, min(tile_v, c - s), 0);
}

// tiles of consecutive voxels (models without a direction)
__global__ void k_tiles_linear(long long n, int tile_v, int4 *tiles, int n_tiles, int *totals)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) totals[0] = n_tiles;
    if (t >= n_tiles) return;
    long long s = (long long)t * tile_v;
    tiles[t] = make_int4(0, (int)s, (int)min((long long)tile_v, n - s), 0);
}

// ------------------------------------------------------------------------------------------------
// Per-direction slab  S[d][r][k], k < n: columns = rotated blocks, optional all-ones "dot" column, isotropic block.
// rot0/rot1: float32 [n0|n1][ndirs][m] (reference layout), iso: float32 [n_iso][m].
__global__ void k_build_slab(const float *__restrict__ rot0, int n0, const float *__restrict__ rot1, int n1, int with_dot,
                             const float *__restrict__ iso, int n_iso, int ndirs, int m, int n_pad, size_t slab_stride,
                             float *slab)
{
    int d = blockIdx.x;
    float *S = slab + (size_t)d * slab_stride;
    int n = n0 + n1 + (with_dot ? 1 : 0) + n_iso;
    #pragma unroll 1
    for (int e = threadIdx.x; e < n * m; e += blockDim.x) {
        int k = e / m, r = e - k * m;
        float v;
        if (k < n0) v = rot0[((size_t)k * ndirs + d) * m + r];
        else if (k < n0 + n1) v = rot1[((size_t)(k - n0) * ndirs + d) * m + r];
        else if (with_dot && k == n0 + n1) v = 1.0f;
        else v = iso[(size_t)(k - n0 - n1 - (with_dot ? 1 : 0)) * m + r];
        S[(size_t)r * n_pad + k] = v;
    }
    // zero the padding columns so that no uninitialised value is ever staged
    #pragma unroll 1
    for (int e = threadIdx.x; e < (n_pad - n) * m; e += blockDim.x) {
        int r = e / (n_pad - n), k = n + e % (n_pad - n);
        S[(size_t)r * n_pad + k] = 0.0f;
    }
}

// SANDI: column-major double (m x n) -> row-major [r][k] double slab
__global__ void k_build_slab_f64(const double *__restrict__ A, int m, int n, int n_pad, double *slab)
{
    #pragma unroll 1
    for (int e = threadIdx.x; e < m * n_pad; e += blockDim.x) {
        int r = e / n_pad, k = e - r * n_pad;
        slab[e] = k < n ? A[(size_t)k * m + r] : 0.0;
    }
}

// Gram of the slab rows listed in `rows` (NULL = all m rows), optionally column-scaled by `norms`
// (norms[jj*ldn + k], jj = position in `rows`): G[i][j] = sum_r (a_ri s_ri)(a_rj s_rj), accumulated
// in row order with un-fused multiply-add exactly like oracle gram_column().
template <typename TS>
__global__ void k_gram(const TS *__restrict__ slab, size_t slab_stride, int n_pad, int K, const int *__restrict__ rows,
                       int nrows, const double *__restrict__ norms, int ldn, int norms_const, double *G, int ldG,
                       size_t G_stride)
{
    int d = blockIdx.x;
    const TS *S = slab + (size_t)d * slab_stride;
    double *Gd = G + (size_t)d * G_stride;
    #pragma unroll 1
    for (int e = threadIdx.x + blockIdx.y * blockDim.x; e < K * K; e += blockDim.x * gridDim.y) {
        int i = e / K, j = e - i * K;
        if (j < i) continue;
        double s = 0.0;
        #pragma unroll 1
        for (int rr = 0; rr < nrows; ++rr) {
            int r = rows ? rows[rr] : rr;
            double ai = (double)S[(size_t)r * n_pad + i], aj = (double)S[(size_t)r * n_pad + j];
            if (norms) {
                int nr = norms_const ? 0 : rr;
                ai = __dmul_rn(ai, norms[(size_t)nr * ldn + i]);
                aj = __dmul_rn(aj, norms[(size_t)nr * ldn + j]);
            }
            s = madd(s, ai, aj);
        }
        Gd[(size_t)i * ldG + j] = s;
        Gd[(size_t)j * ldG + i] = s;
    }
}

// LARS works on G + ridge I: keep the pristine diagonal and write fl(G_kk + ridge) into the table (the same single
// rounding the CPU path applies to its Gram column), so the solver loops carry no ridge special case.
__global__ void k_save_diag(const double *G, int K, int ldG, size_t G_stride, double *diag0)
{
    int d = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x) diag0[(size_t)d * K + k] = G[(size_t)d * G_stride + (size_t)k * ldG + k];
}
__global__ void k_set_ridge(double *G, int K, int ldG, size_t G_stride, const double *diag0, double ridge)
{
    int d = blockIdx.x;
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        G[(size_t)d * G_stride + (size_t)k * ldG + k] = __dadd_rn(diag0[(size_t)d * K + k], ridge);
}

// W[d] = (G_d + ridge I)^-1 for n <= 32 (dense-start NNQP, warp_nnqp_dense): one warp per direction, Gauss-Jordan on [H | I] in
// shared memory, result symmetrised.
__global__ void __launch_bounds__(32) k_invert_spd(const double *__restrict__ H, int n, int ldH, size_t H_stride, double *W, int ldW, size_t W_stride)
{
    __shared__ double M[32][65];
    const int d = blockIdx.x, lane = threadIdx.x;
    const double *Hd = H + (size_t)d * H_stride;
    if (lane < n) {
        for (int q = 0; q < n; ++q) { M[lane][q] = Hd[(size_t)lane * ldH + q]; M[lane][n + q] = (q == lane) ? 1.0 : 0.0; }
    }
    __syncwarp();
    for (int pv = 0; pv < n; ++pv) {
        const double ip = 1.0 / M[pv][pv];
        __syncwarp();
        if (lane == pv) for (int q = 0; q < 2 * n; ++q) M[pv][q] *= ip;
        __syncwarp();
        if (lane < n && lane != pv) {
            const double f = M[lane][pv];
            for (int q = 0; q < 2 * n; ++q) M[lane][q] = fma(-f, M[pv][q], M[lane][q]);
        }
        __syncwarp();
    }
    double *Wd = W + (size_t)d * W_stride;
    if (lane < n)
        for (int q = 0; q < n; ++q) Wd[(size_t)lane * ldW + q] = 0.5 * (M[lane][n + q] + M[q][n + lane]);
}

// ------------------------------------------------------------------------------------------------
struct FitParams {
    int model, m, n, n_pad, ndirs, n_maps, NA;
    // slab
    const void *slab; size_t slab_stride;  // elements per direction
    unsigned slab_bytes;                   // bytes staged per direction (multiple of 16); 0 = read from global
    // Gram tables
    const double *T1; int ldT1; size_t T1_stride;  // NODDI: full dictionary (NNLS stages)
    const double *T2; int ldT2; size_t T2_stride; int K2;  // LARS system
    const double *W; int ldW; size_t W_stride;  // (T2)^-1 per direction: dense-start NNQP of CylinderZeppelinBall (NULL: off)
    // voxels
    const void *y; int y_f64; long long n_vox;
    const int *order; const int4 *tiles; const int *n_tiles_ptr; int *tile_counter;
    double lambda1, lambda2; unsigned flags;
    // NODDI
    const int *dwi_rows; int dc; const double *norms; int norms_const; const float *icvf; const float *kappa; int exvivo;
    int n_wm;
    // FreeWater / CZB / SANDI
    int mouse, n_perp, n_iso, n_rs, n_in;
    const double *Rs, *sandi_norms, *d_in, *d_isos;
    // outputs
    double *est, *rmse, *nrmse, *extra, *coeff_out; int *support_out; long long *status;
    // launch geometry
    int nwarps; unsigned ws_doubles; unsigned slab_smem_off, ws_smem_off;
    int cap_stage[3]; unsigned ws_doubles_stage[3];  // stage kernels: per-stage active-set capacity and workspace size
    int m_pad, dc_pad;
    int fast_lars;    // NODDI stage 2: throughput-oriented LARS (same path, fused arithmetic)
    int aspace;       // NODDI NNLS stages: A-space re-evaluation of near-dependent candidate columns
    int compact3;     // NODDI stage 3: NNLS on the compact support system (one atom per lane) when the support fits a warp
    double *scratch;  // batched NODDI path: per-warp [2][8][NA] doubles
    int batched;
    double *xiso;        // split NODDI path: [n_vox][2] (x_iso, x_dot) by sorted position
    double *c1_all;      // split NODDI path: c1 = A^T y of every voxel [n_vox][NA] by sorted position, written by stage 1 and reused by
                         // stage 3 (NULL: stage 3 recomputes it on the tensor pipe -- volumes whose c1 would not fit the budget)
    const int *lut;      // LUT index per voxel (k_lut)
    int *ovf_list;       // voxels whose active set outgrew a warp: re-fitted by the scalar slow path (amx_slow.cuh)
    long long ovf_cap;
    unsigned *supmask;   // split NODDI path: [n_vox][NPL] stage-2 support, word s bit l <-> atom l + 32 s
    int *exact_list;     // NODDI: exact-fit voxels queued by stage 1 for the A-space QR path (amx_exact.cuh); status[4] counts them
    long long exact_cap;
    double exact_tol;    // ... when ||y - Ax||^2 < exact_tol ||y||^2
    int4 *redo_tiles;    // NODDI stage 3 (thread per voxel): voxels handed back to the warp-per-voxel kernel as one-voxel tiles
    int *redo_count;     // stage 1: [0] their number, [1] the queue head the second pass pulls from; stage 3: [2], [3]
    int tpv3;            // stage 3 runs as k_noddi_stage3_tpv + a second pass over redo_tiles
    int tpv1;            // stage 1 runs as k_noddi_stage1_tpv + a second pass over redo_tiles
};

struct WarpWS {
    double *c1, *dtr, *x, *mat, *rd, *u, *gs, *bx, *y, *y2;
    int *P;
};

constexpr int BV = 8;  // voxels per DMMA micro-batch (the M of m8n8k4)
// alias: the stage kernels never need c1 and dtr at the same time -> one array
// cap: active-set capacity the workspace is laid out for (<= LC); the stage kernels size it per stage -- NODDI stage 1 never
// holds more than ~8 passive atoms -- because every KB of shared memory given back is L1 for the Gram rows
__host__ __device__ inline unsigned ws_doubles_for(int NA, int m_pad, int dc_pad, int alias = 0, int cap = LC)
{
    return (alias == 2 ? 1u : alias ? 2u : 3u) * NA + (unsigned)(cap * (cap + 1) / 2) + 3u * cap + (unsigned)((cap + 1) / 2) + 3 * BV + m_pad + dc_pad;
}

__device__ __forceinline__ WarpWS carve(double *base, int NA, int m_pad, int dc_pad, int alias = 0, int cap = LC)
{
    WarpWS w;
    w.c1 = base; base += NA;
    w.dtr = alias ? w.c1 : base; base += alias ? 0 : NA;
    w.x = alias == 2 ? w.c1 : base; base += alias == 2 ? 0 : NA;  // alias 2 (NODDI stage 2): no coefficient vector at all
    w.mat = base; base += cap * (cap + 1) / 2;
    w.rd = base; base += cap;
    w.u = base; base += cap;
    w.gs = base; base += cap;
    w.P = (int *)base; base += (cap + 1) / 2;
    w.bx = base; base += 3 * BV;  // per-batch x_iso, x_dot, ||y2||^2
    w.y = base; base += m_pad;
    w.y2 = base;
    return w;
}

// out[lane+32s] = sum_r S[row(r)][lane+32s] * scale * yv[r], sequential in r like the CPU loop.  Columns beyond n read
// whatever follows in the slab (in bounds: the slab is padded by NA elements) and produce values nobody uses.
//   FUSED : products are exact in fp64 (fp32-valued dictionary times fp32-valued signal), so fma == mul + add bit for bit;
//   SCALED: NODDI stage 2, column-normalised DWI rows: a = fl(S * norm) first, exactly like the reference's A2.
template <int NPL, typename TS, bool FUSED, bool SCALED>
__device__ __noinline__ double at_y(const TS *S, int n_pad, int n, int nrows, const int *__restrict__ rows, const double *yv,
                                    const double *__restrict__ norms, int ldn, int norms_const, double *out, int lane)
{
    double acc[NPL], nk[NPL];
    double ysq = 0.0;  // sum_r yv[r]^2 in index order, un-fused: the ||y||^2 the LARS stopping rule starts from (rides along for free:
                       // an independent dependency chain next to the accumulators)
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        acc[s] = 0.0;
        nk[s] = 1.0;
        if (SCALED && norms_const && lane + 32 * s < n) nk[s] = norms[lane + 32 * s];
    }
#pragma unroll 4
    for (int rr = 0; rr < nrows; ++rr) {
        const int r = SCALED ? rows[rr] : rr;
        const double yr = yv[rr];
        ysq = madd(ysq, yr, yr);
        const TS *row = S + (size_t)r * n_pad + lane;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            double a = (double)row[32 * s];
            if (SCALED) {
                double sc = nk[s];
                if (!norms_const) sc = (lane + 32 * s < n) ? norms[(size_t)rr * ldn + lane + 32 * s] : 0.0;
                a = __dmul_rn(a, sc);
            }
            acc[s] = FUSED ? fma(a, yr, acc[s]) : madd(acc[s], a, yr);
        }
    }
#pragma unroll
    for (int s = 0; s < NPL; ++s) out[lane + 32 * s] = acc[s];
    __syncwarp();
    return ysq;
}

// sum_i v[i]^2 in index order (all lanes compute the same value)
__device__ __forceinline__ double seq_sumsq(const double *v, int n)
{
    double s = 0.0;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) s = madd(s, v[i], v[i]);
    return s;
}

// Visit the strictly positive entries of x[0..n) in increasing index order; f(j, xj) runs uniformly on all lanes.
template <int NPL, typename F>
__device__ __forceinline__ void for_each_positive(const double *x, int n, int lane, F f)
{
#pragma unroll 1
    for (int s = 0; s < NPL; ++s) {
        int j = lane + 32 * s;
        unsigned mask = __ballot_sync(FULL, j < n && x[j] > 0.0);
        #pragma unroll 1
        while (mask) {
            int l = __ffs(mask) - 1;
            mask &= mask - 1;
            int jj = l + 32 * s;
            f(jj, x[jj]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core micro-GEMM: C[8 voxels][8 NT atoms] = Y[8][rows] * A_d[rows][atoms] with mma.sync m8n8k4 (DMMA).
// The eight voxels of a batch share the direction (same tile), so A_d^T Y is a genuine dense contraction.
// Fragment layout (PTX ISA, m8n8k4 .f64): A row-major: lane holds A[lane/4][lane%4]; B: lane holds B[k=lane%4][n=lane/4];
// C: lane holds C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// c1[v][k] = sum_r y_v[r] A[r][k] for the batch's voxels -> out[v * NA + k]   (NODDI stage 1 / stage 3 right-hand side)
// normy (optional): ||y_v||^2 of the batch's voxels -> normy[v]
template <int NT, int TP, typename TS>
__device__ AMX_GEMM_INLINE void gemm_c1(const TS *S, int n_pad, int m, const void *y, int y_f64, long long myvox, bool vvalid,
                                     double *out, int NA, int lane, double *normy = nullptr)
{
    const int kk = lane & 3, g = lane >> 2;
    const float *yf = (const float *)y + myvox * m;
    const double *yd = (const double *)y + myvox * m;
#pragma unroll 1
    for (int t0 = 0; t0 < NT; t0 += TP) {
        double acc[TP][2];
#pragma unroll
        for (int t = 0; t < TP; ++t) acc[t][0] = acc[t][1] = 0.0;
        double ny = 0.0;
#pragma unroll 1
        for (int r0 = 0; r0 < m; r0 += 4) {
            const int r = r0 + kk;
            const bool rv = r < m;
            double a = 0.0;
            if (rv && vvalid) a = y_f64 ? ld_stream(yd + r) : (double)ld_stream(yf + r);
            ny = fma(a, a, ny);
            const TS *row = S + (size_t)(rv ? r : m - 1) * n_pad + g + 8 * t0;
#pragma unroll
            for (int t = 0; t < TP; ++t) dmma(acc[t][0], acc[t][1], a, (double)ld_stream(row + 8 * t));
        }
        if (normy && t0 == 0) {
            ny += __shfl_xor_sync(FULL, ny, 1);
            ny += __shfl_xor_sync(FULL, ny, 2);
            if (kk == 0) normy[g] = ny;
        }
        double *o = out + (size_t)g * NA + 2 * kk + 8 * t0;
        if (vvalid) {
#pragma unroll
            for (int t = 0; t < TP; ++t) *reinterpret_cast<double2 *>(o + 8 * t) = make_double2(acc[t][0], acc[t][1]);
        }
    }
    __syncwarp();
}

// NODDI stage 2 right-hand side for the batch: y2 = max(y_R - x_iso iso_R [- x_dot], 0) (amico/models.pyx:918-925),
// c2[v][k] = sum_j (A[R_j][k] norms[j][k]) y2_v[j]; also ||y2_v||^2 -> normx[v].
template <int NT, int TP, typename TS, bool NC>
__device__ AMX_GEMM_INLINE void gemm_c2(const TS *S, int n_pad, int n, int n_wm, int dc, const int *__restrict__ rows, const void *y,
                                     int y_f64, int m, long long myvox, bool vvalid, double xiso, double xdot, int exvivo,
                                     const double *__restrict__ norms, int ldn, double *out, int NA, double *normx, int lane)
{
    const int kk = lane & 3, g = lane >> 2;
    const float *yf = (const float *)y + myvox * m;
    const double *yd = (const double *)y + myvox * m;
#pragma unroll 1
    for (int t0 = 0; t0 < NT; t0 += TP) {
        double acc[TP][2];
#pragma unroll
        for (int t = 0; t < TP; ++t) acc[t][0] = acc[t][1] = 0.0;
        double nx = 0.0;
#pragma unroll 1
        for (int j0 = 0; j0 < dc; j0 += 4) {
            const int jj = j0 + kk;
            const bool jv = jj < dc;
            const int r = rows[jv ? jj : dc - 1];
            const TS *row = S + (size_t)r * n_pad;
            double a = 0.0;
            if (jv && vvalid) {
                a = (y_f64 ? ld_stream(yd + r) : (double)ld_stream(yf + r)) - xiso * (double)row[n - 1];
                if (exvivo) a = a - xdot * 1.0;
                a = a < 0.0 ? 0.0 : a;
            }
            nx = fma(a, a, nx);
            row += g + 8 * t0;
            if (NC) {
#pragma unroll
                for (int t = 0; t < TP; ++t) dmma(acc[t][0], acc[t][1], a, (double)ld_stream(row + 8 * t));
            } else {
                const double *nr = norms + (size_t)(jv ? jj : dc - 1) * ldn + g + 8 * t0;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    const double sc = (g + 8 * (t0 + t) < n_wm) ? nr[8 * t] : 0.0;
                    dmma(acc[t][0], acc[t][1], a, __dmul_rn((double)ld_stream(row + 8 * t), sc));
                }
            }
        }
        if (t0 == 0) {
            nx += __shfl_xor_sync(FULL, nx, 1);
            nx += __shfl_xor_sync(FULL, nx, 2);
            if (kk == 0) normx[g] = nx;  // g < BV always
        }
        double *o = out + (size_t)g * NA + 2 * kk + 8 * t0;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            double c0 = acc[t][0], c1 = acc[t][1];
            if (NC) {
                const int k0 = 8 * (t0 + t) + 2 * kk;
                c0 = (k0 < n_wm) ? c0 * norms[k0] : 0.0;
                c1 = (k0 + 1 < n_wm) ? c1 * norms[k0 + 1] : 0.0;
            }
            if (vvalid) *reinterpret_cast<double2 *>(o + 8 * t) = make_double2(c0, c1);
        }
    }
    __syncwarp();
}

// NODDI maps from the final coefficients (amico/models.pyx:945-979); all lanes compute, lane 0 stores
template <int NPL>
__device__ __noinline__ void noddi_maps(const float *__restrict__ icvf, const float *__restrict__ kappa, int n, int n_wm, int exvivo,
                                        unsigned flags, double *e, double *emod, const double *x, int lane)
{
    double s_all = 0.0;
    for_each_positive<NPL>(x, n, lane, [&](int, double xj) { s_all += xj; });
    s_all += 1e-16;
    double s_wm = 0.0;
    for_each_positive<NPL>(x, n_wm, lane, [&](int, double xj) { s_wm += xj / s_all; });
    s_wm += 1e-16;
    double f1 = 0.0, f2 = 0.0, k1 = 0.0;
    for_each_positive<NPL>(x, n_wm, lane, [&](int j, double xj) {
        float ic = icvf[j];
        f1 += (double)ic * xj / s_all / s_wm;
        f2 += (double)((float)(1.0 - (double)ic)) * xj / s_all / s_wm;
        k1 += (double)kappa[j] * xj / s_all / s_wm;
    });
    const double ndi = f1 / (f1 + f2 + 1e-16);
    const double odi = 2.0 / 3.14159265358979323846 * atan2(1.0, k1);
    const double fwf = x[n - 1] / s_all;
    if (lane == 0) {
        e[0] = ndi; e[1] = odi; e[2] = fwf;
        if (exvivo) e[3] = x[n - 2] / s_all;
        if (flags & FLAG_EXTRA) {
            double tf = 1.0 - fwf;
            emod[0] = ndi * tf;
            emod[1] = odi * tf;
        }
    }
}

// fit errors (amico/models.pyx:45-71): y_est = A x with the full dictionary; ws.y is overwritten
template <int NPL, typename TS>
__device__ __noinline__ void fit_errors(const TS *S, int n_pad, int n, int m, double *yv, const double *x, unsigned flags,
                                           double *rmse_out, double *nrmse_out, int lane)
{
    double den = 0.0;
    if (flags & FLAG_NRMSE) den = seq_sumsq(yv, m);
    __syncwarp();
    #pragma unroll 1
    for (int i = lane; i < m; i += 32) {
        double ye = 0.0;
        #pragma unroll 1
        for (int j = 0; j < n; ++j) {
            double xj = x[j];
            if (xj != 0.0) ye = madd(ye, (double)S[(size_t)i * n_pad + j], xj);
        }
        double d = yv[i] - ye;
        yv[i] = d * d;
    }
    __syncwarp();
    if (flags & FLAG_RMSE) {
        double acc = 0.0;
        #pragma unroll 1
        for (int i = 0; i < m; ++i) acc += yv[i] / (double)m;
        if (lane == 0) *rmse_out = sqrt(acc);
    }
    if (flags & FLAG_NRMSE) {
        double acc = 0.0;
        if (den > 1e-16) {
            #pragma unroll 1
            for (int i = 0; i < m; ++i) acc += yv[i] / den;
            acc = sqrt(acc);
        }
        if (lane == 0) *nrmse_out = acc;
    }
}

// Maps of the single-fit models from the elastic-net coefficients x (amico/models.pyx:1241-1255 FreeWater, :618-636
// CylinderZeppelinBall, :1570-1611 SANDI -- x is un-normalised in place there); all lanes compute, lane 0 stores.  Returns the
// support size.
template <int MODEL, int NPL>
__device__ __forceinline__ int lasso_maps(const FitParams &p, double *x, long long vox, int lane)
{
    const int n = p.n;
    int support = 0;
    for_each_positive<NPL>(x, n, lane, [&](int, double) { ++support; });
    if (MODEL == MODEL_FREEWATER) {
        double xs = 0.0, xp = 0.0;
        for_each_positive<NPL>(x, n, lane, [&](int j, double xj) { xs += xj; if (j < p.n_perp) xp += xj; });
        xs += 1e-16;
        const double vv = xp / xs;
        if (lane == 0) {
            double *e = p.est + vox * p.n_maps;
            e[0] = vv; e[1] = 1.0 - vv;
            if (p.mouse) { e[2] = x[p.n_perp] / xs; e[3] = x[p.n_perp + 1] / xs; }
        }
    } else if (MODEL == MODEL_CZB) {
        double f1 = 0.0, f2 = 0.0, aa = 0.0;
        for_each_positive<NPL>(x, p.n_rs + p.n_perp, lane, [&](int j, double xj) { if (j < p.n_rs) f1 += xj; else f2 += xj; });
        f2 += 1e-16;
        const double vv = f1 / (f1 + f2 + 1e-16);
        f1 += 1e-16;
        for_each_positive<NPL>(x, p.n_rs, lane, [&](int j, double xj) { aa += p.Rs[j] * xj; });
        aa = 1e6 * 2.0 * aa / f1;
        const double dd = (4.0 * vv) / (3.14159265358979323846 * (aa * aa) + 1e-16);
        if (lane == 0) {
            double *e = p.est + vox * 3;
            e[0] = vv; e[1] = aa; e[2] = dd;
        }
    } else {  // SANDI: un-normalise, then group sums
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int j = lane + 32 * s;
            if (j < n) x[j] = x[j] * p.sandi_norms[j];
        }
        __syncwarp();
        const int n_rs = p.n_rs, n_in = p.n_in;
        double xs = 0, sph = 0, stk = 0, iso = 0, Rsoma = 0, Din = 0, De = 0;
        for_each_positive<NPL>(x, n, lane, [&](int j, double xj) {
            xs += xj;
            if (j < n_rs) sph += xj;
            else if (j < n_rs + n_in) stk += xj;
            else iso += xj;
        });
        xs += 1e-16;
        for_each_positive<NPL>(x, n, lane, [&](int j, double xj) {
            if (j < n_rs) Rsoma += p.Rs[j] * xj;
            else if (j < n_rs + n_in) Din += p.d_in[j - n_rs] * xj;
            else De += p.d_isos[j - n_rs - n_in] * xj;
        });
        if (lane == 0) {
            double *e = p.est + vox * 6;
            e[0] = sph / xs; e[1] = stk / xs; e[2] = iso / xs;
            sph += 1e-16; stk += 1e-16; iso += 1e-16;
            e[3] = 1e6 * Rsoma / sph; e[4] = 1e3 * Din / stk; e[5] = 1e3 * De / iso;
        }
    }
    return support;
}

// Optional outputs of the single-fit models: support / coefficients (debug), FreeWater corrected DWI (:1263-1274), fit errors
// (:45-71).  yv: the voxel's signal as doubles in shared memory (destroyed by the error pass).
template <int MODEL, int NPL, typename TS>
__device__ __forceinline__ void lasso_outputs(const FitParams &p, const TS *S, double *yv, const double *x, long long vox, int support, int lane)
{
    const int m = p.m, n = p.n, n_pad = p.n_pad;
    if (p.support_out && lane == 0) p.support_out[vox] = support;
    if (p.coeff_out)
        #pragma unroll 1
        for (int j = lane; j < n; j += 32) p.coeff_out[vox * n + j] = x[j];
    if (MODEL == MODEL_FREEWATER && (p.flags & FLAG_EXTRA)) {
        #pragma unroll 1
        for (int i = lane; i < m; i += 32) {
            double fw = 0.0;
            #pragma unroll 1
            for (int k = n - p.n_iso; k < n; ++k) fw = madd(fw, (double)S[(size_t)i * n_pad + k], x[k]);
            double cv = yv[i] - fw;
            p.extra[vox * m + i] = cv < 0.0 ? 0.0 : cv;
        }
        __syncwarp();
    }
    if (p.flags & (FLAG_RMSE | FLAG_NRMSE))
        fit_errors<NPL, TS>(S, n_pad, n, m, yv, x, p.flags, p.rmse ? p.rmse + vox : nullptr, p.nrmse ? p.nrmse + vox : nullptr, lane);
}

// queue a voxel for the scalar slow path (status[2] counts the entries)
__device__ __forceinline__ void queue_slow(const FitParams &p, long long vox, int lane)
{
    if (lane == 0) {
        unsigned long long idx = atomicAdd((unsigned long long *)&p.status[2], 1ull);
        if ((long long)idx < p.ovf_cap) p.ovf_list[idx] = (int)vox;
    }
}

template <int MODEL, int NPL, typename TS>
__global__ void __launch_bounds__(512, 1) k_fit(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *mbar = (uint64_t *)smem;
    int *s_tile = (int *)(smem + 8);
    int *s_next = (int *)(smem + 12);
    TS *s_slab = (TS *)(smem + p.slab_smem_off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpWS ws = carve((double *)(smem + p.ws_smem_off) + (size_t)warp * p.ws_doubles, p.NA, p.m_pad, p.dc_pad);
    const int m = p.m, n = p.n, n_pad = p.n_pad;
    const bool staged = p.slab_bytes != 0;
    uint32_t phase = 0;
    if (threadIdx.x == 0 && staged) {
        mbar_init(mbar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    long long n_overflow = 0;

    for (;;) {
        if (threadIdx.x == 0) *s_tile = atomicAdd(p.tile_counter, 1);
        __syncthreads();
        const int t = *s_tile;
        if (t >= *p.n_tiles_ptr) break;
        const int4 tile = p.tiles[t];
        const int dir = tile.x;
        const TS *Sg = (const TS *)p.slab + (size_t)dir * p.slab_stride;
        const TS *S = Sg;
        if (staged) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(mbar, p.slab_bytes);
                bulk_g2s(s_slab, Sg, p.slab_bytes, mbar);
                *s_next = p.nwarps;
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            S = s_slab;
        } else if (threadIdx.x == 0) {
            *s_next = p.nwarps;
        }
        __syncthreads();
        const double *T1 = p.T1 ? p.T1 + (size_t)dir * p.T1_stride : nullptr;
        const double *T2 = p.T2 + (size_t)dir * p.T2_stride;

        int v = warp;
        while (v < tile.z) {
            const long long vox = p.order ? (long long)p.order[tile.y + v] : (long long)tile.y + v;
            // ---- signal
            if (p.y_f64) {
                const double *yg = (const double *)p.y + vox * m;
                #pragma unroll 1
                for (int i = lane; i < m; i += 32) ws.y[i] = yg[i];
            } else {
                const float *yg = (const float *)p.y + vox * m;
                #pragma unroll 1
                for (int i = lane; i < m; i += 32) ws.y[i] = (double)yg[i];
            }
            __syncwarp();
            int overflow = 0, support = 0;

            if (MODEL == MODEL_NODDI) {
                const int n_wm = p.n_wm;
                // stage 1: isotropic fraction (amico/models.pyx:911)
                if (p.y_f64) at_y<NPL, TS, false, false>(S, n_pad, n, m, nullptr, ws.y, nullptr, 0, 0, ws.c1, lane);
                else at_y<NPL, TS, true, false>(S, n_pad, n, m, nullptr, ws.y, nullptr, 0, 0, ws.c1, lane);
                unsigned all = 0;
#pragma unroll
                for (int s = 0; s < NPL; ++s) all |= (lane + 32 * s < n ? 1u : 0u) << s;
                overflow |= warp_nnls<NPL>(T1, p.ldT1, n, m, 3 * n, ws.c1, ws.x, all, ws.mat, ws.rd, ws.P, lane, nullptr);
                const double xiso = ws.x[n - 1];
                const double xdot = p.exvivo ? ws.x[n - 2] : 0.0;
                __syncwarp();
                // stage 2: support selection on the normalised DWI rows (:914-926)
                #pragma unroll 1
                for (int jj = lane; jj < p.dc; jj += 32) {
                    int r = p.dwi_rows[jj];
                    double v2 = ws.y[r] - xiso * (double)S[(size_t)r * n_pad + (n - 1)];
                    if (p.exvivo) v2 = v2 - xdot * 1.0;
                    ws.y2[jj] = v2 < 0.0 ? 0.0 : v2;
                }
                __syncwarp();
                const double normX = at_y<NPL, TS, false, true>(S, n_pad, n_wm, p.dc, p.dwi_rows, ws.y2, p.norms, n_wm, p.norms_const, ws.dtr, lane);
                overflow |= warp_lars<NPL>(T2, p.ldT2, p.lambda2, n_wm, p.dc < n_wm ? p.dc : n_wm, p.lambda1, ws.dtr, normX,
                                           ws.mat, ws.u, ws.gs, ws.P, ws.x, lane, nullptr);
                // stage 3: debias on the support (:929-942)
                unsigned allowed = 0;
#pragma unroll
                for (int s = 0; s < NPL; ++s) {
                    int j = lane + 32 * s;
                    bool on = (j < n_wm && ws.x[j] > 0.0) || (j >= n_wm && j < n);
                    allowed |= (on ? 1u : 0u) << s;
                    support += __popc(__ballot_sync(FULL, on));
                }
                __syncwarp();
                overflow |= warp_nnls<NPL>(T1, p.ldT1, n, m, 3 * support, ws.c1, ws.x, allowed, ws.mat, ws.rd, ws.P, lane, nullptr);
                noddi_maps<NPL>(p.icvf, p.kappa, n, n_wm, p.exvivo, p.flags, p.est + vox * p.n_maps,
                                (p.flags & FLAG_EXTRA) ? p.extra + 2 * vox : nullptr, ws.x, lane);
            } else if (MODEL != MODEL_NODDI) {
                // single elastic-net fit on the full dictionary (:615, :1238, :1569)
                const double normX = (p.y_f64 || sizeof(TS) == 8)
                                         ? at_y<NPL, TS, false, false>(S, n_pad, n, m, nullptr, ws.y, nullptr, 0, 0, ws.dtr, lane)
                                         : at_y<NPL, TS, true, false>(S, n_pad, n, m, nullptr, ws.y, nullptr, 0, 0, ws.dtr, lane);
                overflow |= warp_lars<NPL>(T2, p.ldT2, p.lambda2, n, m < n ? m : n, p.lambda1, ws.dtr, normX, ws.mat, ws.u,
                                           ws.gs, ws.P, ws.x, lane, nullptr);
                support = lasso_maps<MODEL, NPL>(p, ws.x, vox, lane);
            }
            __syncwarp();
            if (MODEL == MODEL_NODDI) {
                if (p.support_out && lane == 0) p.support_out[vox] = support;
                if (p.coeff_out)
                    #pragma unroll 1
                    for (int j = lane; j < n; j += 32) p.coeff_out[vox * n + j] = ws.x[j];
                if (p.flags & (FLAG_RMSE | FLAG_NRMSE))
                    fit_errors<NPL, TS>(S, n_pad, n, m, ws.y, ws.x, p.flags, p.rmse ? p.rmse + vox : nullptr,
                                        p.nrmse ? p.nrmse + vox : nullptr, lane);
            } else {
                lasso_outputs<MODEL, NPL, TS>(p, S, ws.y, ws.x, vox, support, lane);
            }
            if (overflow) {
                if (MODEL == MODEL_NODDI && p.ovf_list) queue_slow(p, vox, lane);  // re-fitted by the scalar slow path
                else ++n_overflow;
            }
            __syncwarp();
            // next voxel of this tile
            int nv = 0;
            if (lane == 0) nv = atomicAdd(s_next, 1);
            v = __shfl_sync(FULL, nv, 0);
        }
        __syncthreads();
    }
    if (lane == 0 && n_overflow) atomicAdd((unsigned long long *)&p.status[3], (unsigned long long)n_overflow);
}

// ------------------------------------------------------------------------------------------------
// FreeWater / CylinderZeppelinBall / SANDI, throughput path: the NODDI stage-2 design applied to the single-fit models.  Every warp
// pulls 8-voxel batches of one LUT direction from a global queue; c = A^T y of the batch runs on the FP64 tensor pipe (DMMA
// m8n8k4, M = the 8 voxels; ||y||^2 rides along), then each voxel walks the SPAMS homotopy path with warp_lars_fast (same path and
// stopping rules as the reference's lasso -- amico/models.pyx:615, 1238, 1569 --, fused arithmetic) and its maps are formed in the
// same warp.  The elastic net is strictly convex, so the minimiser is unique and the maps agree with the bit-exact kernel (k_fit,
// AMX_FLAG_EXACT) to ~1e-12; that one follows the reference's un-fused CPU arithmetic operation for operation and stays available.
template <int MODEL, int NPL, typename TS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_lasso_batched(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = p.cap_stage[1];
    WarpWS ws = carve((double *)(smem + p.ws_smem_off) + (size_t)warp * p.ws_doubles_stage[1], p.NA, p.m_pad, 0, 1, cap);
    constexpr int NT = 4 * NPL, TP = (MAXT > 768 && NPL > 1) ? NT / 2 : NT;
    const int m = p.m, n = p.n, n_pad = p.n_pad, NA = p.NA;
    double *scr = p.scratch + ((size_t)blockIdx.x * 32 + warp) * (size_t)BV * NA;
    const int g = lane >> 2;
    const int n_tiles = *p.n_tiles_ptr;
    long long n_overflow = 0;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(p.tile_counter, 1);
        b = __shfl_sync(FULL, b, 0);
        if (b >= n_tiles) break;
        const int4 tile = p.tiles[b];
        const int nb = tile.z;  // <= BV
        const TS *S = (const TS *)p.slab + (size_t)tile.x * p.slab_stride;
        const double *T2 = p.T2 + (size_t)tile.x * p.T2_stride;
        const bool vvalid = g < nb;
        const long long mypos = tile.y + (vvalid ? g : 0);
        const long long myvox = p.order ? (long long)p.order[mypos] : mypos;
        gemm_c1<NT, TP, TS>(S, n_pad, m, p.y, p.y_f64, myvox, vvalid, scr, NA, lane, ws.bx);
        #pragma unroll 1
        for (int v = 0; v < nb; ++v) {
            const long long vox = p.order ? (long long)p.order[tile.y + v] : (long long)tile.y + v;
#pragma unroll
            for (int s = 0; s < NPL; ++s) ws.dtr[lane + 32 * s] = scr[(size_t)v * NA + lane + 32 * s];
            __syncwarp();
            bool solved = false;
            if (MODEL == MODEL_CZB && NPL == 1 && p.W) {  // dense minimiser: start from the unconstrained ridge solution (certified)
                double xl = 0.0;
                solved = warp_nnqp_dense(T2, p.ldT2, p.W + (size_t)tile.x * p.W_stride, p.ldW, n, ws.dtr[lane], ws.mat, xl, lane);
                if (solved) ws.x[lane] = xl;
                __syncwarp();
            }
            if (!solved) {
                unsigned sup[NPL];
                const int ov = warp_lars_fast<NPL>(T2, p.ldT2, n, m < n ? m : n, p.lambda1, ws.dtr, ws.bx[v], ws.mat, ws.u, ws.gs, ws.P, ws.x,
                                                   lane, cap, sup);
                if (ov) ++n_overflow;
            }
            const int support = lasso_maps<MODEL, NPL>(p, ws.x, vox, lane);
            __syncwarp();
            if (p.m_pad) {  // debug / error / corrected-signal outputs want the signal as doubles
                if (p.y_f64) {
                    const double *yg = (const double *)p.y + vox * m;
                    #pragma unroll 1
                    for (int i = lane; i < m; i += 32) ws.y[i] = yg[i];
                } else {
                    const float *yg = (const float *)p.y + vox * m;
                    #pragma unroll 1
                    for (int i = lane; i < m; i += 32) ws.y[i] = (double)yg[i];
                }
                __syncwarp();
                lasso_outputs<MODEL, NPL, TS>(p, S, ws.y, ws.x, vox, support, lane);
            } else if (p.support_out || p.coeff_out) {
                if (p.support_out && lane == 0) p.support_out[vox] = support;
                if (p.coeff_out)
                    for (int j = lane; j < n; j += 32) p.coeff_out[vox * n + j] = ws.x[j];
            }
            __syncwarp();
        }
    }
    if (lane == 0 && n_overflow) atomicAdd((unsigned long long *)&p.status[3], (unsigned long long)n_overflow);
}

// ------------------------------------------------------------------------------------------------
// NODDI as three stage kernels over the same batch queue (STAGE 1: NNLS for the isotropic fraction, 2: LARS support
// selection, 3: NNLS on the support + maps).  One solver per kernel: a fused kernel's
// fused kernel's ~90 KB of SASS thrashes the instruction cache once 16 independent warps per SM sit in different
// solvers (ncu: `no_instruction` was the top stall).  Each stage kernel keeps one solver hot.  Between stages only
// 16 B (x_iso, x_dot) + 4 NPL B (support mask) per voxel travel through HBM; stage 3 recomputes c1 = A^T y on the
// tensor pipe instead of storing 1.2 KB per voxel.
template <int STAGE, int NPL, typename TS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_noddi_stage(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = p.cap_stage[STAGE - 1];
    WarpWS ws = carve((double *)(smem + p.ws_smem_off) + (size_t)warp * p.ws_doubles_stage[STAGE - 1], p.NA, p.m_pad, p.dc_pad,
                      (STAGE == 2 && p.fast_lars) ? 2 : 1, cap);
    constexpr int NT = 4 * NPL, TP = (MAXT > 512 && NT % 2 == 0) ? NT / 2 : NT;
    const int m = p.m, n = p.n, n_pad = p.n_pad, n_wm = p.n_wm, NA = p.NA;
    double *scr = p.scratch + ((size_t)blockIdx.x * 32 + warp) * (size_t)BV * NA;  // 32 = most warps any stage launches
    const int g = lane >> 2;
    unsigned all = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) all |= (lane + 32 * s < n ? 1u : 0u) << s;
    int *counter = p.tile_counter + (STAGE - 1);
    const int n_tiles = *p.n_tiles_ptr;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1);
        b = __shfl_sync(FULL, b, 0);
        if (b >= n_tiles) break;
        const int4 tile = p.tiles[b];
        const int nb = tile.z;  // <= BV
        const TS *S = (const TS *)p.slab + (size_t)tile.x * p.slab_stride;
        const bool vvalid = g < nb;
        const long long mypos = tile.y + (vvalid ? g : 0);
        const long long myvox = (long long)p.order[mypos];
        if (STAGE == 2) {
            const double *T2 = p.T2 + (size_t)tile.x * p.T2_stride;
            const double xi = p.xiso[2 * mypos], xd = p.xiso[2 * mypos + 1];
            if (p.norms_const)
                gemm_c2<NT, TP, TS, true>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo, p.norms,
                                      n_wm, scr, NA, ws.bx, lane);
            else
                gemm_c2<NT, TP, TS, false>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo, p.norms,
                                       n_wm, scr, NA, ws.bx, lane);
            #pragma unroll 1
            for (int v = 0; v < nb; ++v) {
#pragma unroll
                for (int s = 0; s < NPL; ++s) ws.dtr[lane + 32 * s] = scr[(size_t)v * NA + lane + 32 * s];
                __syncwarp();
                int ov;
                if (p.fast_lars) {  // the support comes back as bit words: no coefficient vector in shared memory
                    unsigned sup[NPL];
                    ov = warp_lars_fast<NPL>(T2, p.ldT2, n_wm, p.dc < n_wm ? p.dc : n_wm, p.lambda1, ws.dtr, ws.bx[v], ws.mat, ws.u, ws.gs,
                                             ws.P, nullptr, lane, cap, sup);
#pragma unroll
                    for (int s = 0; s < NPL; ++s) {
                        const int j = lane + 32 * s;
                        const unsigned w = sup[s] | __ballot_sync(FULL, j >= n_wm && j < n);  // dot / iso columns always belong
                        if (lane == s) p.supmask[(size_t)(tile.y + v) * NPL + s] = w;
                    }
                } else {
                    ov = warp_lars<NPL>(T2, p.ldT2, p.lambda2, n_wm, p.dc < n_wm ? p.dc : n_wm, p.lambda1, ws.dtr, ws.bx[v], ws.mat, ws.u,
                                        ws.gs, ws.P, ws.x, lane, nullptr, cap);
#pragma unroll
                    for (int s = 0; s < NPL; ++s) {
                        const int j = lane + 32 * s;
                        const unsigned w = __ballot_sync(FULL, (j < n_wm && ws.x[j] > 0.0) || (j >= n_wm && j < n));
                        if (lane == s) p.supmask[(size_t)(tile.y + v) * NPL + s] = w;
                    }
                }
                if (ov) queue_slow(p, (long long)p.order[tile.y + v], lane);
                __syncwarp();
            }
        } else {
            const double *T1 = p.T1 + (size_t)tile.x * p.T1_stride;
            // c1 of the batch: stage 1 computes it (DMMA) -- into the per-voxel store when there is one, so that stage 3 only reads it
            double *c1b = p.c1_all ? p.c1_all + (size_t)tile.y * NA : scr;
            if (STAGE == 1 || !p.c1_all) gemm_c1<NT, TP, TS>(S, n_pad, m, p.y, p.y_f64, myvox, vvalid, c1b, NA, lane, STAGE == 1 ? ws.bx : nullptr);
            #pragma unroll 1
            for (int v = 0; v < nb; ++v) {
                const long long pos = tile.y + v;
#pragma unroll
                for (int s = 0; s < NPL; ++s) ws.c1[lane + 32 * s] = c1b[(size_t)v * NA + lane + 32 * s];
                __syncwarp();
                const ASpace asp{(const float *)S, n_pad, m, p.y, p.y_f64, (long long)p.order[pos]};
                const ASpace *as = (p.aspace && sizeof(TS) == 4) ? &asp : nullptr;
                if (STAGE == 1) {  // isotropic fraction (amico/models.pyx:911)
                    double zz = 0.0;
                    int ov = warp_nnls<NPL>(T1, p.ldT1, n, m, 3 * n, ws.c1, ws.x, all, ws.mat, ws.rd, ws.P, lane, nullptr, cap, nullptr, as, &zz);
                    if (lane == 0) {
                        p.xiso[2 * pos] = ws.x[n - 1];
                        p.xiso[2 * pos + 1] = p.exvivo ? ws.x[n - 2] : 0.0;
                        // exact-fit voxel (residual below exact_tol ||y||^2): the Gram-space pivots are decided by rounding noise there --
                        // queue it for the A-space QR path, which re-fits it from scratch after stage 3
                        const double yy = ws.bx[v];
                        if (yy > 0.0 && yy - zz < p.exact_tol * yy) {
                            const unsigned long long idx = atomicAdd((unsigned long long *)&p.status[4], 1ull);
                            if ((long long)idx < p.exact_cap) p.exact_list[idx] = (int)p.order[pos];
                        }
                    }
                    if (ov) queue_slow(p, (long long)p.order[pos], lane);
                } else {  // debias on the support (:929-942), maps
                    const long long vox = (long long)p.order[pos];
                    unsigned allowed = 0;
                    int support = 0;
#pragma unroll
                    for (int s = 0; s < NPL; ++s) {
                        const unsigned w = p.supmask[(size_t)pos * NPL + s];
                        allowed |= ((w >> lane) & 1u) << s;
                        support += __popc(w);
                    }
                    int ov;
                    const double *xf = ws.x;  // coefficients by atom index
                    if (support <= 32 && p.compact3) {
                        // compact system: lane q owns the q-th atom of the support (ascending atom order)
                        int *map = (int *)ws.u;
                        int q = lane, atom = 0;
#pragma unroll
                        for (int s = 0; s < NPL; ++s) {
                            const unsigned w = p.supmask[(size_t)pos * NPL + s];
                            const int cnt = __popc(w);
                            if (q >= 0 && q < cnt) { atom = 32 * s + (int)__fns(w, 0, q + 1); q = -1; }
                            else if (q >= 0) q -= cnt;
                        }
                        const double cq = lane < support ? ws.c1[atom] : 0.0;
                        __syncwarp();
                        map[lane] = atom;
                        ws.c1[lane] = cq;  // c of the compact system, in place (every lane has read its entry)
                        __syncwarp();
                        ov = warp_nnls<1, true>(T1, p.ldT1, support, m, 3 * support, ws.c1, ws.x, lane < support ? 1u : 0u, ws.mat, ws.rd,
                                                ws.P, lane, nullptr, cap, map, as);
                        const double xq = lane < support ? ws.x[lane] : 0.0;
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < NPL; ++s) ws.c1[lane + 32 * s] = 0.0;
                        __syncwarp();
                        if (lane < support) ws.c1[atom] = xq;
                        __syncwarp();
                        xf = ws.c1;
                    } else {
                        ov = warp_nnls<NPL>(T1, p.ldT1, n, m, 3 * support, ws.c1, ws.x, allowed, ws.mat, ws.rd, ws.P, lane, nullptr, cap, nullptr, as);
                    }
                    noddi_maps<NPL>(p.icvf, p.kappa, n, n_wm, p.exvivo, p.flags, p.est + vox * p.n_maps,
                                    (p.flags & FLAG_EXTRA) ? p.extra + 2 * vox : nullptr, xf, lane);
                    if (p.support_out && lane == 0) p.support_out[vox] = support;
                    if (p.coeff_out)
                        for (int j = lane; j < n; j += 32) p.coeff_out[vox * n + j] = xf[j];
                    if (p.flags & (FLAG_RMSE | FLAG_NRMSE)) {
                        if (p.y_f64) {
                            const double *yg = (const double *)p.y + vox * m;
                            #pragma unroll 1
                            for (int i = lane; i < m; i += 32) ws.y[i] = yg[i];
                        } else {
                            const float *yg = (const float *)p.y + vox * m;
                            #pragma unroll 1
                            for (int i = lane; i < m; i += 32) ws.y[i] = (double)yg[i];
                        }
                        __syncwarp();
                        fit_errors<NPL, TS>(S, n_pad, n, m, ws.y, xf, p.flags, p.rmse ? p.rmse + vox : nullptr,
                                            p.nrmse ? p.nrmse + vox : nullptr, lane);
                    }
                    if (ov) queue_slow(p, vox, lane);
                }
                __syncwarp();
            }
        }
    }
}

}  // namespace amx
