// Exact path of the NODDI fit: the reference's own algorithm -- Lawson-Hanson NNLS on A with Householder QR
// (amico/models.pyx:911, 940 -> cyspams nnls), statement for statement as restated in oracle/amico_oracle.c::nnls_core --
// for the voxels the Gram-space stage kernels cannot resolve: EXACT-FIT voxels (noise-free phantoms, ||y - Ax||^2 below
// 1e-6 ||y||^2).  There the passive Gram system has condition ~1e14 and the dual c - Hx is rounding noise; the stage-1
// kernel detects the regime from ||y||^2 - ||z||^2 and queues the voxel, and this kernel re-fits it from scratch.
//
// One CTA per voxel.  Column-parallel parts of the algorithm (dual vector, Householder application to the zero-set
// columns) run thread per column with the CPU's summation order; everything sequential runs on thread 0 in the CPU's
// order, and all arithmetic is un-fused (-fmad=false): the result is bit-identical to the oracle's, like the lasso models.
// The LARS stage in between is warp_lars (bit-exact restatement of the SPAMS path) on warp 0.
#pragma once
#include "amx_kernels.cuh"

namespace amx {

// Householder "H12" of Lawson & Hanson (oracle/amico_oracle.c::h12_construct / h12_apply)
__device__ inline void h12_construct(int lp, int l1, int m, double *u, double *up)
{
    double cl = fabs(u[lp]), sm, clinv;
    if (l1 >= m) { *up = 0.0; return; }
    for (int j = l1; j < m; ++j) if (fabs(u[j]) > cl) cl = fabs(u[j]);
    if (cl <= 0.0) { *up = 0.0; return; }
    clinv = 1.0 / cl;
    sm = (u[lp] * clinv) * (u[lp] * clinv);
    for (int j = l1; j < m; ++j) sm += (u[j] * clinv) * (u[j] * clinv);
    cl *= sqrt(sm);
    if (u[lp] > 0.0) cl = -cl;
    *up = u[lp] - cl;
    u[lp] = cl;
}

__device__ inline void h12_apply(int lp, int l1, int m, const double *u, double up, double *c)
{
    double b = up * u[lp], sm;
    if (l1 >= m) return;
    if (b >= 0.0) return;
    b = 1.0 / b;
    sm = c[lp] * up;
    for (int i = l1; i < m; ++i) sm += c[i] * u[i];
    if (sm != 0.0) {
        sm *= b;
        c[lp] += sm * up;
        for (int i = l1; i < m; ++i) c[i] += sm * u[i];
    }
}

__device__ inline void lh_givens(double a, double b, double *c, double *s, double *sig)
{
    double xr, yr;
    if (fabs(a) > fabs(b)) {
        xr = b / a; yr = sqrt(1.0 + xr * xr);
        *c = copysign(1.0 / yr, a); *s = (*c) * xr; *sig = fabs(a) * yr;
    } else if (b != 0.0) {
        xr = a / b; yr = sqrt(1.0 + xr * xr);
        *s = copysign(1.0 / yr, b); *c = (*s) * xr; *sig = fabs(b) * yr;
    } else { *sig = 0.0; *c = 0.0; *s = 1.0; }
}

struct ExactWS {
    double *a;            // global: m x n working copy of the dictionary, column-major (destroyed)
    double *b, *zz;       // shared: m
    double *w, *x;        // shared: n
    int *index;           // shared: n
    int *ctl;             // shared: [0] iz1 [1] nsetp [2] npp1 [3] iter [4] j [5] state (0 run, 1 done) ; double up in dctl[0]
    double *dctl;
};

// min ||a x - y||, x >= 0 over the n columns of `a` (oracle/amico_oracle.c::nnls_core).  All threads of the CTA call it.
__device__ inline void nnls_cta(const ExactWS &W, const double *y, int m, int n, double *x)
{
    double *a = W.a, *b = W.b, *zz = W.zz, *w = W.w;
    int *index = W.index, *ctl = W.ctl;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int iz2 = n - 1, itmax = 3 * n;
#define AE(r, c) a[(size_t)(c) * m + (r)]
    for (int i = tid; i < n; i += nt) { x[i] = 0.0; index[i] = i; }
    for (int i = tid; i < m; i += nt) b[i] = y[i];
    if (tid == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; ctl[3] = 0; ctl[4] = 0; ctl[5] = 0; }
    __syncthreads();
    for (;;) {
        const int iz1 = ctl[0], nsetp0 = ctl[1], npp1_0 = ctl[2];
        if (!(iz1 <= iz2 && nsetp0 < m)) break;
        // dual vector on the zero set: thread per column, rows in order
        for (int iz = iz1 + tid; iz <= iz2; iz += nt) {
            const int j = index[iz];
            double sm = 0.0;
            for (int l = npp1_0; l < m; ++l) sm += AE(l, j) * b[l];
            w[j] = sm;
        }
        __syncthreads();
        if (tid == 0) {
            int npp1 = npp1_0, j = 0, iz = 0, izmax = 0;
            double up = 0.0;
            bool found = false;
            for (;;) {
                double wmax = 0.0;
                for (iz = iz1; iz <= iz2; ++iz) {
                    j = index[iz];
                    if (w[j] > wmax) { wmax = w[j]; izmax = iz; }
                }
                if (wmax <= 0.0) break;
                iz = izmax; j = index[iz];
                const double asave = AE(npp1, j);
                h12_construct(npp1, npp1 + 1, m, &AE(0, j), &up);
                double unorm = 0.0;
                for (int l = 0; l < nsetp0; ++l) unorm += AE(l, j) * AE(l, j);
                unorm = sqrt(unorm);
                const double tmp = unorm + fabs(AE(npp1, j)) * 0.01;
                if (tmp - unorm > 0.0) {
                    for (int l = 0; l < m; ++l) zz[l] = b[l];
                    h12_apply(npp1, npp1 + 1, m, &AE(0, j), up, zz);
                    const double ztest = zz[npp1] / AE(npp1, j);
                    if (ztest > 0.0) { found = true; break; }
                }
                AE(npp1, j) = asave;
                w[j] = 0.0;
            }
            if (found) {
                for (int l = 0; l < m; ++l) b[l] = zz[l];
                index[iz] = index[iz1]; index[iz1] = j;
                ctl[0] = iz1 + 1; ctl[1] = npp1 + 1; ctl[2] = npp1 + 1; ctl[4] = j;
                W.dctl[0] = up;
            } else {
                ctl[5] = 1;
            }
        }
        __syncthreads();
        if (ctl[5]) break;
        {   // apply the new reflector to the remaining zero-set columns: thread per column
            const int nsetp = ctl[1], npp1 = ctl[2], j = ctl[4], iz1n = ctl[0];
            const double up = W.dctl[0];
            for (int jz = iz1n + tid; jz <= iz2; jz += nt) h12_apply(nsetp - 1, npp1, m, &AE(0, j), up, &AE(0, index[jz]));
        }
        __syncthreads();
        if (tid == 0) {
            int iz1n = ctl[0], nsetp = ctl[1], npp1 = ctl[2], iter = ctl[3], j = ctl[4], jj = 0, ip, ii, i, l;
            bool hit_cap = false;
            for (l = npp1; l < m; ++l) AE(l, j) = 0.0;
            w[j] = 0.0;
            for (l = 0; l < m; ++l) zz[l] = b[l];
            for (ip = nsetp - 1; ip >= 0; --ip) {
                if (ip != nsetp - 1) for (ii = 0; ii <= ip; ++ii) zz[ii] -= AE(ii, jj) * zz[ip + 1];
                jj = index[ip];
                zz[ip] /= AE(ip, jj);
            }
            for (;;) {  // secondary loop
                if (++iter > itmax) { hit_cap = true; break; }
                double alpha = 2.0;
                jj = -1;
                for (ip = 0; ip < nsetp; ++ip) {
                    l = index[ip];
                    if (zz[ip] <= 0.0) {
                        const double t = -x[l] / (zz[ip] - x[l]);
                        if (alpha > t) { alpha = t; jj = ip; }
                    }
                }
                if (alpha == 2.0) break;
                for (ip = 0; ip < nsetp; ++ip) { l = index[ip]; x[l] += alpha * (zz[ip] - x[l]); }
                i = index[jj];
                for (;;) {
                    x[i] = 0.0;
                    if (jj != nsetp - 1) {
                        ++jj;
                        for (j = jj; j < nsetp; ++j) {
                            double cc, ss, sig, tmp;
                            ii = index[j]; index[j - 1] = ii;
                            lh_givens(AE(j - 1, ii), AE(j, ii), &cc, &ss, &sig);
                            AE(j - 1, ii) = sig; AE(j, ii) = 0.0;
                            for (l = 0; l < n; ++l) if (l != ii) {
                                tmp = AE(j - 1, l);
                                AE(j - 1, l) = cc * tmp + ss * AE(j, l);
                                AE(j, l) = -ss * tmp + cc * AE(j, l);
                            }
                            tmp = b[j - 1];
                            b[j - 1] = cc * tmp + ss * b[j];
                            b[j] = -ss * tmp + cc * b[j];
                        }
                    }
                    npp1 = nsetp - 1; --nsetp; --iz1n; index[iz1n] = i;
                    for (jj = 0; jj < nsetp; ++jj) { i = index[jj]; if (x[i] <= 0.0) break; }
                    if (jj == nsetp) break;
                }
                for (l = 0; l < m; ++l) zz[l] = b[l];
                for (ip = nsetp - 1; ip >= 0; --ip) {
                    if (ip != nsetp - 1) for (ii = 0; ii <= ip; ++ii) zz[ii] -= AE(ii, jj) * zz[ip + 1];
                    jj = index[ip];
                    zz[ip] /= AE(ip, jj);
                }
            }
            if (hit_cap) {
                ctl[5] = 1;
            } else {
                for (ip = 0; ip < nsetp; ++ip) x[index[ip]] = zz[ip];
            }
            ctl[0] = iz1n; ctl[1] = nsetp; ctl[2] = npp1; ctl[3] = iter;
        }
        __syncthreads();
        if (ctl[5]) break;
    }
#undef AE
    __syncthreads();
}

// shared-memory doubles the exact kernel needs next to warp 0's solver workspace
__host__ __device__ inline unsigned exact_extra_doubles(int m, int NA) { return 2u * ((m + 1) & ~1) + 3u * NA + 16u; }

template <int NPL, typename TS>
__global__ void __launch_bounds__(32 * NPL) k_noddi_exact(const FitParams p, const int *__restrict__ list, long long *status, double *scratch_a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const long long count = min(status[4], p.exact_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) status[7] += count;  // call total (status[4] is per voxel chunk)
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m, n = p.n, n_pad = p.n_pad, n_wm = p.n_wm, dc = p.dc, NA = p.NA;
    const int m_pad = (m + 1) & ~1, dc_pad = (dc + 1) & ~1;
    double *base = (double *)smem;
    WarpWS ws = carve(base, NA, m_pad, dc_pad, 0, LC);
    double *ex = base + ws_doubles_for(NA, m_pad, dc_pad, 0, LC);
    ExactWS W;
    W.a = scratch_a + (size_t)blockIdx.x * ((size_t)m * n);
    W.b = ex; ex += m_pad;
    W.zz = ex; ex += m_pad;
    W.w = ex; ex += NA;
    W.x = ex; ex += NA;                 // stage-3 coefficients in compact numbering
    W.index = (int *)ex; ex += NA / 2;
    int *pos = (int *)ex; ex += NA / 2;
    W.dctl = ex; ex += 2;
    W.ctl = (int *)ex;
    int *s_pc = W.ctl + 8;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long vox = list[it];
        const int dir = p.lut[vox];
        const TS *S = (const TS *)p.slab + (size_t)dir * p.slab_stride;
        const double *T2 = p.T2 + (size_t)dir * p.T2_stride;
        // signal and the full dictionary (amico/models.pyx:905-908), float64 column-major like the reference's A
        for (int i = tid; i < m; i += nt) ws.y[i] = p.y_f64 ? ((const double *)p.y)[vox * m + i] : (double)((const float *)p.y)[vox * m + i];
        for (int e = tid; e < m * n; e += nt) {
            const int r = e / n, k = e - r * n;
            W.a[(size_t)k * m + r] = (double)S[(size_t)r * n_pad + k];
        }
        __syncthreads();
        // fit 1: isotropic fraction (:911)
        nnls_cta(W, ws.y, m, n, ws.x);
        int overflow = 0;
        if (warp == 0) {
            // fit 2: support selection on the normalised DWI rows (:914-926)
            const double xiso = ws.x[n - 1], xdot = p.exvivo ? ws.x[n - 2] : 0.0;
            __syncwarp();
            #pragma unroll 1
            for (int jj = lane; jj < dc; jj += 32) {
                const int r = p.dwi_rows[jj];
                double v2 = ws.y[r] - xiso * (double)S[(size_t)r * n_pad + (n - 1)];
                if (p.exvivo) v2 = v2 - xdot * 1.0;
                ws.y2[jj] = v2 < 0.0 ? 0.0 : v2;
            }
            __syncwarp();
            const double normX = at_y<NPL, TS, false, true>(S, n_pad, n_wm, dc, p.dwi_rows, ws.y2, p.norms, n_wm, p.norms_const, ws.dtr, lane);
            overflow = warp_lars<NPL>(T2, p.ldT2, p.lambda2, n_wm, dc < n_wm ? dc : n_wm, p.lambda1, ws.dtr, normX, ws.mat, ws.u, ws.gs,
                                      ws.P, ws.x, lane, nullptr);
            // fit 3 operands: the support plus the dot / isotropic columns (:929-939)
            if (lane == 0) {
                int pc = 0;
                for (int j = 0; j < n; ++j)
                    if (j >= n_wm || ws.x[j] > 0.0) pos[pc++] = j;
                *s_pc = pc;
            }
            __syncwarp();
        }
        __syncthreads();
        const int pc = *s_pc;
        for (int e = tid; e < m * pc; e += nt) {
            const int r = e / pc, k = e - r * pc;
            W.a[(size_t)k * m + r] = (double)S[(size_t)r * n_pad + pos[k]];
        }
        __syncthreads();
        nnls_cta(W, ws.y, m, pc, W.x);
        for (int j = tid; j < NA; j += nt) ws.x[j] = 0.0;
        __syncthreads();
        for (int j = tid; j < pc; j += nt) ws.x[pos[j]] = W.x[j];
        __syncthreads();
        if (warp == 0) {
            noddi_maps<NPL>(p.icvf, p.kappa, n, n_wm, p.exvivo, p.flags, p.est + vox * p.n_maps,
                            (p.flags & FLAG_EXTRA) ? p.extra + 2 * vox : nullptr, ws.x, lane);
            if (p.support_out && lane == 0) p.support_out[vox] = pc;
            if (p.coeff_out)
                for (int j = lane; j < n; j += 32) p.coeff_out[vox * n + j] = ws.x[j];
            if (p.flags & (FLAG_RMSE | FLAG_NRMSE))
                fit_errors<NPL, TS>(S, n_pad, n, m, ws.y, ws.x, p.flags, p.rmse ? p.rmse + vox : nullptr, p.nrmse ? p.nrmse + vox : nullptr, lane);
            if (overflow) queue_slow(p, vox, lane);  // support beyond a warp (lambda1 far below default): the scalar path takes over
        }
        __syncthreads();
    }
}

}  // namespace amx
