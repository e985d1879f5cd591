// Scalar slow path: one THREAD per voxel, workspace in global memory, no limit on the active-set size.
//
// The warp solvers keep one active atom per lane (<= 32).  With the reference's default regularisation that is ample
// (NODDI supports: mean 10, max 25 of 145), but a user may lower lambda1 (set_solver) and get supports of 50-90 atoms.
// Voxels whose active set would outgrow a warp are queued by the fast kernels and re-fitted here from scratch with the
// same algorithms (Lawson-Hanson pivoting on the Gram system with a Cholesky factor; the SPAMS LARS path of
// oracle/amico_oracle.c::lars_core) written as plain sequential code.  Throughput is irrelevant here; correctness is not.
#pragma once
#include "amx_kernels.cuh"

namespace amx {

struct SlowWS {
    double *L;      // cap x cap (Cholesky factor, or LARS inverse, upper stored like the CPU code)
    double *z, *s, *g, *v, *coef;  // cap
    double *w, *c1, *c2, *x, *dtr, *slope, *tq;  // NA
    int *P;         // cap
    unsigned char *inP, *allowed;  // NA
};

__host__ __device__ inline size_t slow_ws_bytes(int cap, int NA)
{
    size_t b = (size_t)cap * cap * 8 + (size_t)5 * cap * 8 + (size_t)7 * NA * 8 + (size_t)cap * 4 + 2 * (size_t)NA;
    return (b + 255) & ~(size_t)255;
}

__device__ inline SlowWS slow_carve(unsigned char *base, int cap, int NA)
{
    SlowWS w;
    double *d = (double *)base;
    w.L = d; d += (size_t)cap * cap;
    w.z = d; d += cap; w.s = d; d += cap; w.g = d; d += cap; w.v = d; d += cap; w.coef = d; d += cap;
    w.w = d; d += NA; w.c1 = d; d += NA; w.c2 = d; d += NA; w.x = d; d += NA; w.dtr = d; d += NA; w.slope = d; d += NA; w.tq = d; d += NA;
    w.P = (int *)d;
    unsigned char *c = (unsigned char *)(w.P + cap);
    w.inP = c; w.allowed = c + NA;
    return w;
}

// Lawson-Hanson NNLS on (T, c) restricted to `allowed`; same rules as warp_nnls.
__device__ inline void slow_nnls(const double *__restrict__ T, int ld, const double *c, int n, int mcap, int itmax, double *x,
                                 const unsigned char *allowed, SlowWS &W, int cap)
{
    double *L = W.L, *z = W.z, *s = W.s, *g = W.g, *v = W.v, *w = W.w;
    int *P = W.P;
    unsigned char *inP = W.inP;
    int np = 0, iter = 0;
    for (int j = 0; j < n; ++j) { x[j] = 0.0; inP[j] = 0; }
    while (np < mcap && np < cap) {
        for (int j = 0; j < n; ++j) {
            if (inP[j] || !allowed[j]) { w[j] = 0.0; continue; }
            double acc = c[j];
            for (int k = 0; k < np; ++k) acc = fma(-T[(size_t)P[k] * ld + j], x[P[k]], acc);
            w[j] = acc;
        }
        int j = -1;
        double d2 = 0.0, znew = 0.0;
        for (;;) {
            double wmax = 0.0;
            j = -1;
            for (int k = 0; k < n; ++k)
                if (!inP[k] && allowed[k] && w[k] > wmax) { wmax = w[k]; j = k; }
            if (j < 0) break;
            double vv = 0.0, vz = 0.0;
            for (int a = 0; a < np; ++a) g[a] = T[(size_t)P[a] * ld + j];
            for (int k = 0; k < np; ++k) {  // forward substitution, column oriented like the warp code
                double vk = g[k] / L[(size_t)k * cap + k];
                v[k] = vk;
                for (int a = k + 1; a < np; ++a) g[a] = fma(-L[(size_t)a * cap + k], vk, g[a]);
            }
            for (int a = 0; a < np; ++a) { vv = fma(v[a], v[a], vv); vz = fma(v[a], z[a], vz); }
            d2 = T[(size_t)j * ld + j] - vv;
            bool ok = false;
            if (d2 > 0.0) {
                double unorm = sqrt(vv), dd = sqrt(d2), tt = unorm + dd * 0.01;
                if (tt - unorm > 0.0) {
                    znew = (c[j] - vz) / dd;
                    ok = znew > 0.0;
                }
            }
            if (ok) break;
            w[j] = 0.0;
        }
        if (j < 0) break;
        for (int a = 0; a < np; ++a) L[(size_t)np * cap + a] = v[a];
        L[(size_t)np * cap + np] = sqrt(d2);
        z[np] = znew;
        P[np++] = j;
        inP[j] = 1;
        bool feasible = false;
        for (;;) {
            if (++iter > itmax) return;
            for (int a = 0; a < np; ++a) s[a] = z[a];
            for (int k = np - 1; k >= 0; --k) {  // back substitution
                double sk = s[k] / L[(size_t)k * cap + k];
                s[k] = sk;
                for (int a = 0; a < k; ++a) s[a] = fma(-L[(size_t)k * cap + a], sk, s[a]);
            }
            double alpha = 2.0;
            int jj = -1;
            for (int a = 0; a < np; ++a)
                if (s[a] <= 0.0) {
                    double t = -x[P[a]] / (s[a] - x[P[a]]);
                    if (t < alpha) { alpha = t; jj = a; }
                }
            if (jj < 0) { feasible = true; break; }
            for (int a = 0; a < np; ++a) x[P[a]] = fma(alpha, s[a] - x[P[a]], x[P[a]]);
            x[P[jj]] = 0.0;
            int q = 0;
            for (int a = 0; a < np; ++a) {
                if (x[P[a]] <= 0.0) { inP[P[a]] = 0; x[P[a]] = 0.0; }
                else P[q++] = P[a];
            }
            np = q;
            if (np == 0) { feasible = true; break; }
            for (int i = 0; i < np; ++i) {  // rebuild the factor and z = L^-1 c_P
                for (int jx = 0; jx <= i; ++jx) {
                    double sm = T[(size_t)P[i] * ld + P[jx]];
                    for (int k = 0; k < jx; ++k) sm = fma(-L[(size_t)i * cap + k], L[(size_t)jx * cap + k], sm);
                    L[(size_t)i * cap + jx] = (i == jx) ? sqrt(sm > 0.0 ? sm : 1e-300) : sm / L[(size_t)jx * cap + jx];
                }
                double sm = c[P[i]];
                for (int k = 0; k < i; ++k) sm = fma(-L[(size_t)i * cap + k], z[k], sm);
                z[i] = sm / L[(size_t)i * cap + i];
            }
        }
        if (feasible)
            for (int a = 0; a < np; ++a) x[P[a]] = s[a];
    }
}

// Non-negative LARS on T (= G + ridge I), DtR destroyed; oracle/amico_oracle.c::lars_core with the Gram table.
__device__ inline void slow_lars(const double *__restrict__ T, int ld, double *DtR, double normX, int K, int Ltrue, double lambda1,
                                 double *x, SlowWS &W, int cap)
{
    double *invGs = W.L, *u = W.v, *gs = W.g, *coeffs = W.coef, *slope = W.slope, *tq = W.tq;
    int *ind = W.P;
    int L = Ltrue < K ? Ltrue : K;
    if (L > cap) L = cap;
    const int LL = cap;
    for (int k = 0; k < K; ++k) x[k] = 0.0;
    if (L <= 0) return;
    for (int j = 0; j < L; ++j) { coeffs[j] = 0.0; ind[j] = -1; }
    int currentInd = 0;
    for (int k = 1; k < K; ++k)
        if (DtR[k] > DtR[currentInd]) currentInd = k;
    if (fabs(DtR[currentInd]) < lambda1) return;
    int newAtom = 1, iter = 0;
    const int length_path = 4 * L;
#define SYMU(r, c) ((r) <= (c) ? invGs[(size_t)(c) * LL + (r)] : invGs[(size_t)(r) * LL + (c)])
    for (int i = 0; i < L; ++i) {
        if (i < 0) break;
        ++iter;
        if (newAtom) {
            ind[i] = currentInd;
            for (int j = 0; j <= i; ++j) gs[j] = T[(size_t)currentInd * ld + ind[j]];
            if (i == 0) invGs[0] = 1.0 / gs[0];
            else {
                double dot = 0.0;
                for (int r = 0; r < i; ++r) {
                    double sm = 0.0;
                    for (int c = 0; c < i; ++c) sm = madd(sm, SYMU(r, c), gs[c]);
                    u[r] = sm;
                }
                for (int j = 0; j < i; ++j) dot = madd(dot, u[j], gs[j]);
                const double schur = 1.0 / __dsub_rn(gs[i], dot);
                invGs[(size_t)i * LL + i] = schur;
                for (int j = 0; j < i; ++j) invGs[(size_t)i * LL + j] = __dmul_rn(-schur, u[j]);
                for (int k = 0; k < i; ++k)
                    for (int j = 0; j <= k; ++j)
                        invGs[(size_t)k * LL + j] = __dadd_rn(invGs[(size_t)k * LL + j], __dmul_rn(__dmul_rn(schur, u[j]), u[k]));
            }
        }
        for (int j = 0; j <= i; ++j) gs[j] = DtR[ind[j]] > 0.0 ? 1.0 : -1.0;
        for (int r = 0; r <= i; ++r) {
            double sm = 0.0;
            for (int c = 0; c <= i; ++c) sm = madd(sm, SYMU(r, c), gs[c]);
            u[r] = sm;
        }
        double step_max = INFINITY;
        int first_zero = -1;
        for (int j = 0; j <= i; ++j) {
            double ratio = -coeffs[j] / u[j];
            if (ratio > 0.0 && ratio <= step_max) { step_max = ratio; first_zero = j; }
        }
        const double cc = fabs(DtR[ind[0]]);
        for (int k = 0; k < K; ++k) {
            double sm = 0.0;
            for (int j = 0; j <= i; ++j) sm = madd(sm, T[(size_t)ind[j] * ld + k], u[j]);
            slope[k] = sm;
            tq[k] = sm;
        }
        for (int j = 0; j <= i; ++j) tq[ind[j]] = INFINITY;
        for (int k = 0; k < K; ++k) tq[k] = (tq[k] < INFINITY && tq[k] < 1.0) ? __ddiv_rn(__dsub_rn(cc, DtR[k]), __dsub_rn(1.0, tq[k])) : INFINITY;
        int index = 0;
        double best = fabs(tq[0]);
        for (int k = 1; k < K; ++k)
            if (fabs(tq[k]) < best) { best = fabs(tq[k]); index = k; }
        double step = tq[index];
        currentInd = index;
        double coeff1 = 0.0, coeff2 = 0.0;
        for (int j = 0; j <= i; ++j) coeff1 = __dadd_rn(coeff1, DtR[ind[j]] > 0.0 ? u[j] : -u[j]);
        for (int j = 0; j <= i; ++j) coeff2 = madd(coeff2, DtR[ind[j]], u[j]);
        const double step_max2 = __dsub_rn(cc, lambda1);
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        for (int j = 0; j <= i; ++j) {
            coeffs[j] = madd(coeffs[j], step, u[j]);
            if (coeffs[j] < 0.0) coeffs[j] = 0.0;
        }
        for (int k = 0; k < K; ++k) DtR[k] = __dsub_rn(DtR[k], __dmul_rn(step, slope[k]));
        normX = __dadd_rn(normX, __dsub_rn(__dmul_rn(__dmul_rn(coeff1, step), step), __dmul_rn(__dmul_rn(2.0, coeff2), step)));
        if (step == step_max) {
            const int zr = first_zero;
            for (int j = zr; j < i; ++j) { ind[j] = ind[j + 1]; coeffs[j] = coeffs[j + 1]; }
            ind[i] = -1; coeffs[i] = 0.0;
            const double schur = invGs[(size_t)zr * LL + zr];
            for (int k = 0; k < zr; ++k) u[k] = invGs[(size_t)zr * LL + k];
            for (int k = zr; k < i; ++k) u[k] = invGs[(size_t)(k + 1) * LL + zr];
            for (int j = zr; j < i; ++j) {
                for (int k = 0; k < zr; ++k) invGs[(size_t)j * LL + k] = invGs[(size_t)(j + 1) * LL + k];
                for (int k = zr; k <= j; ++k) invGs[(size_t)j * LL + k] = invGs[(size_t)(j + 1) * LL + k + 1];
            }
            for (int k = 0; k < i; ++k)
                for (int j = 0; j <= k; ++j)
                    invGs[(size_t)k * LL + j] = __dsub_rn(invGs[(size_t)k * LL + j], __ddiv_rn(__dmul_rn(u[j], u[k]), schur));
            newAtom = 0;
            i -= 2;
        } else {
            newAtom = 1;
        }
        if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == L - 1) break;
    }
#undef SYMU
    for (int j = 0; j < L; ++j)
        if (ind[j] >= 0) x[ind[j]] = coeffs[j];
}

// NODDI, whole pipeline for the queued voxels (amico/models.pyx:901-981)
template <typename TS>
__global__ void k_slow_noddi(const FitParams p, const int *__restrict__ list, long long *status, unsigned char *wsbase,
                             size_t ws_bytes, int cap)
{
    const long long count = min(status[2], p.ovf_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) status[6] += count;  // call total (status[2] is per voxel chunk)
    const int m = p.m, n = p.n, n_pad = p.n_pad, n_wm = p.n_wm, dc = p.dc;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    SlowWS W = slow_carve(wsbase + (size_t)tid * ws_bytes, cap, p.NA);
    for (long long it = tid; it < count; it += nth) {
        const long long vox = list[it];
        const int dir = p.lut[vox];
        const TS *S = (const TS *)p.slab + (size_t)dir * p.slab_stride;
        const double *T1 = p.T1 + (size_t)dir * p.T1_stride;
        const double *T2 = p.T2 + (size_t)dir * p.T2_stride;
        const float *yf = (const float *)p.y + vox * m;
        const double *yd = (const double *)p.y + vox * m;
#define YV(r) (p.y_f64 ? yd[r] : (double)yf[r])
        for (int k = 0; k < n; ++k) {
            double sm = 0.0;
            for (int r = 0; r < m; ++r) sm = fma((double)S[(size_t)r * n_pad + k], YV(r), sm);
            W.c1[k] = sm;
            W.allowed[k] = 1;
        }
        slow_nnls(T1, p.ldT1, W.c1, n, m, 3 * n, W.x, W.allowed, W, cap);
        const double xiso = W.x[n - 1], xdot = p.exvivo ? W.x[n - 2] : 0.0;
        double nx = 0.0;
        for (int k = 0; k < n_wm; ++k) W.c2[k] = 0.0;
        for (int jj = 0; jj < dc; ++jj) {
            const int r = p.dwi_rows[jj];
            double a = YV(r) - xiso * (double)S[(size_t)r * n_pad + n - 1];
            if (p.exvivo) a = a - xdot * 1.0;
            a = a < 0.0 ? 0.0 : a;
            nx = fma(a, a, nx);
            for (int k = 0; k < n_wm; ++k) {
                const double sc = p.norms[(size_t)(p.norms_const ? 0 : jj) * n_wm + k];
                W.c2[k] = fma(__dmul_rn((double)S[(size_t)r * n_pad + k], sc), a, W.c2[k]);
            }
        }
        for (int k = 0; k < n_wm; ++k) W.dtr[k] = W.c2[k];
        slow_lars(T2, p.ldT2, W.dtr, nx, n_wm, dc < n_wm ? dc : n_wm, p.lambda1, W.x, W, cap);
        int support = 0;
        for (int k = 0; k < n; ++k) {
            W.allowed[k] = (k < n_wm) ? (W.x[k] > 0.0) : 1;
            support += W.allowed[k];
        }
        slow_nnls(T1, p.ldT1, W.c1, n, m, 3 * support, W.x, W.allowed, W, cap);
        // maps (:945-979)
        double s_all = 0.0, s_wm = 0.0, f1 = 0.0, f2 = 0.0, k1 = 0.0;
        for (int k = 0; k < n; ++k) s_all += W.x[k];
        s_all += 1e-16;
        for (int k = 0; k < n_wm; ++k) s_wm += W.x[k] / s_all;
        s_wm += 1e-16;
        for (int k = 0; k < n_wm; ++k) {
            const float ic = p.icvf[k];
            f1 += (double)ic * W.x[k] / s_all / s_wm;
            f2 += (double)((float)(1.0 - (double)ic)) * W.x[k] / s_all / s_wm;
            k1 += (double)p.kappa[k] * W.x[k] / s_all / s_wm;
        }
        const double ndi = f1 / (f1 + f2 + 1e-16), odi = 2.0 / 3.14159265358979323846 * atan2(1.0, k1), fwf = W.x[n - 1] / s_all;
        double *e = p.est + vox * p.n_maps;
        e[0] = ndi; e[1] = odi; e[2] = fwf;
        if (p.exvivo) e[3] = W.x[n - 2] / s_all;
        if (p.flags & FLAG_EXTRA) { p.extra[2 * vox] = ndi * (1.0 - fwf); p.extra[2 * vox + 1] = odi * (1.0 - fwf); }
        if (p.support_out) p.support_out[vox] = support;
        if (p.coeff_out)
            for (int k = 0; k < n; ++k) p.coeff_out[vox * n + k] = W.x[k];
        if (p.flags & (FLAG_RMSE | FLAG_NRMSE)) {
            double den = 0.0, acc_r = 0.0, acc_n = 0.0;
            for (int r = 0; r < m; ++r) den = madd(den, YV(r), YV(r));
            for (int r = 0; r < m; ++r) {
                double ye = 0.0;
                for (int k = 0; k < n; ++k)
                    if (W.x[k] != 0.0) ye = madd(ye, (double)S[(size_t)r * n_pad + k], W.x[k]);
                const double dd = YV(r) - ye;
                acc_r += dd * dd / (double)m;
                if (den > 1e-16) acc_n += dd * dd / den;
            }
            if (p.flags & FLAG_RMSE) p.rmse[vox] = sqrt(acc_r);
            if (p.flags & FLAG_NRMSE) p.nrmse[vox] = den > 1e-16 ? sqrt(acc_n) : 0.0;
        }
#undef YV
    }
}

}  // namespace amx
