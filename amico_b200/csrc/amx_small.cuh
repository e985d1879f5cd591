// Thread-per-voxel fit for SMALL dictionaries (n <= 16 atoms): FreeWater (10 zeppelins + 1-2 balls, m = 33) and SANDI (15 atoms,
// m = 1 + n_shells = 4).  A warp solving ONE such voxel leaves most of its lanes idle -- the homotopy path holds <= 11 / <= 4
// active atoms and every step is a handful of flops -- so here a warp solves 32 voxels, one per lane, each walking the same SPAMS
// path (amico/models.pyx:1238, 1569 -> cyspams lasso, mode PENALTY, pos = true; same entering / leaving / stopping rules as
// warp_lars_fast and oracle/amico_oracle.c::lars_core, fused arithmetic) with its state in registers / thread-local arrays.
// Voxels are taken in LUT-direction order, so the lanes of a warp read the same dictionary slab and Gram table (L1 broadcast).
// Bit-reproducible results: AMX_FLAG_EXACT -> k_fit.
#pragma once
#include "amx_kernels.cuh"

namespace amx {

// x[0..n) -> maps of the single-fit models, scalar restatement of lasso_maps (sums over the positive entries in index order).
template <int MODEL, int NMAX>
__device__ __forceinline__ int small_maps(const FitParams &p, double (&x)[NMAX], long long vox)
{
    const int n = p.n;
    int support = 0;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) support += (j < n && x[j] > 0.0) ? 1 : 0;
    if (MODEL == MODEL_FREEWATER) {
        double xs = 0.0, xp = 0.0;
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n && x[j] > 0.0) { xs += x[j]; if (j < p.n_perp) xp += x[j]; }
        xs += 1e-16;
        const double vv = xp / xs;
        double *e = p.est + vox * p.n_maps;
        e[0] = vv; e[1] = 1.0 - vv;
        if (p.mouse) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) { if (j == p.n_perp) a = x[j]; if (j == p.n_perp + 1) b = x[j]; }
            e[2] = a / xs; e[3] = b / xs;
        }
    } else {  // SANDI (amico/models.pyx:1570-1611): un-normalise, then group sums
        const int n_rs = p.n_rs, n_in = p.n_in;
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n) x[j] = x[j] * p.sandi_norms[j];
        double xs = 0, sph = 0, stk = 0, iso = 0, Rsoma = 0, Din = 0, De = 0;
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n && x[j] > 0.0) {
                xs += x[j];
                if (j < n_rs) sph += x[j];
                else if (j < n_rs + n_in) stk += x[j];
                else iso += x[j];
            }
        xs += 1e-16;
#pragma unroll
        for (int j = 0; j < NMAX; ++j)
            if (j < n && x[j] > 0.0) {
                if (j < n_rs) Rsoma += p.Rs[j] * x[j];
                else if (j < n_rs + n_in) Din += p.d_in[j - n_rs] * x[j];
                else De += p.d_isos[j - n_rs - n_in] * x[j];
            }
        double *e = p.est + vox * 6;
        e[0] = sph / xs; e[1] = stk / xs; e[2] = iso / xs;
        sph += 1e-16; stk += 1e-16; iso += 1e-16;
        e[3] = 1e6 * Rsoma / sph; e[4] = 1e3 * Din / stk; e[5] = 1e3 * De / iso;
    }
    return support;
}

// One voxel per thread.  NMAX: atom capacity (registers), LMAX: active-set capacity = min(m, n) of the model.
template <int MODEL, typename TS, int NMAX, int LMAX>
__global__ void __launch_bounds__(128) k_lasso_small(const FitParams p)
{
    const int m = p.m, n = p.n, n_pad = p.n_pad, ldT = p.ldT2;
    const int L = min(min(m, n), LMAX);
    const double lambda1 = p.lambda1;
    for (long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x; pos < p.n_vox; pos += (long long)gridDim.x * blockDim.x) {
        const long long vox = p.order ? (long long)p.order[pos] : pos;
        const int dir = p.lut ? p.lut[vox] : 0;
        const TS *S = (const TS *)p.slab + (size_t)dir * p.slab_stride;
        const double *T = p.T2 + (size_t)dir * p.T2_stride;
        const float *yf = (const float *)p.y + vox * m;
        const double *yd = (const double *)p.y + vox * m;
        // correlations c = A^T y and ||y||^2
        double DtR[NMAX], x[NMAX];
#pragma unroll
        for (int k = 0; k < NMAX; ++k) { DtR[k] = 0.0; x[k] = 0.0; }
        double normX = 0.0;
        #pragma unroll 1
        for (int r = 0; r < m; ++r) {
            const double yr = p.y_f64 ? yd[r] : (double)yf[r];
            normX = fma(yr, yr, normX);
            const TS *row = S + (size_t)r * n_pad;
#pragma unroll
            for (int k = 0; k < NMAX; ++k)
                if (k < n) DtR[k] = fma((double)row[k], yr, DtR[k]);
        }
        // ---- homotopy path
        double Mi[LMAX * LMAX];  // (G_SS)^-1, full square, symmetric
        double coef[LMAX], u[LMAX], gs[LMAX];
        int ind[LMAX];
#pragma unroll
        for (int j = 0; j < LMAX; ++j) { coef[j] = 0.0; ind[j] = -1; u[j] = 0.0; gs[j] = 0.0; }
        int cur = 0;
        {
            double bv = DtR[0];
#pragma unroll
            for (int k = 1; k < NMAX; ++k)
                if (k < n && DtR[k] > bv) { bv = DtR[k]; cur = k; }
            if (!(fabs(bv) < lambda1) && L > 0) {
                int newAtom = 1, iter = 0, na = 0;
                unsigned act = 0;
                const int length_path = 4 * L;
                #pragma unroll 1
                for (int i = 0; i < L; ++i) {
                    if (i < 0) break;
                    ++iter;
                    if (newAtom) {
                        ind[i] = cur;
                        coef[i] = 0.0;
                        act |= 1u << cur;
                        const double *Tc = T + (size_t)cur * ldT;
                        #pragma unroll 1
                        for (int j = 0; j <= i; ++j) gs[j] = Tc[ind[j]];
                        if (i == 0) {
                            Mi[0] = 1.0 / gs[0];
                        } else {
                            double dot = 0.0;
                            #pragma unroll 1
                            for (int r = 0; r < i; ++r) {
                                double s = 0.0;
                                #pragma unroll 1
                                for (int c = 0; c < i; ++c) s = fma(Mi[r * LMAX + c], gs[c], s);
                                u[r] = s;
                                dot = fma(s, gs[r], dot);
                            }
                            const double schur = 1.0 / (gs[i] - dot);
                            #pragma unroll 1
                            for (int r = 0; r < i; ++r) {
                                const double su = schur * u[r];
                                #pragma unroll 1
                                for (int c = 0; c < i; ++c) Mi[r * LMAX + c] = fma(su, u[c], Mi[r * LMAX + c]);
                                Mi[r * LMAX + i] = -su;
                                Mi[i * LMAX + r] = -su;
                            }
                            Mi[i * LMAX + i] = schur;
                        }
                    }
                    na = i + 1;
                    // path direction u = invGs * sign(DtR_S)
                    #pragma unroll 1
                    for (int j = 0; j <= i; ++j) gs[j] = DtR[ind[j]] > 0.0 ? 1.0 : -1.0;
                    #pragma unroll 1
                    for (int r = 0; r <= i; ++r) {
                        double s = 0.0;
                        #pragma unroll 1
                        for (int c = 0; c <= i; ++c) s = fma(Mi[r * LMAX + c], gs[c], s);
                        u[r] = s;
                    }
                    // largest step before an active coefficient crosses zero (last index wins ties)
                    double step_max = INFINITY;
                    int fz = -1;
                    #pragma unroll 1
                    for (int j = 0; j <= i; ++j) {
                        const double r = -coef[j] / u[j];
                        if (r > 0.0 && r <= step_max) { step_max = r; fz = j; }
                    }
                    const double cc = fabs(DtR[ind[0]]);
                    // correlation slopes and the first inactive atom reaching the common correlation (smallest |step|, lowest index)
                    double sl[NMAX];
                    double best = INFINITY, step = INFINITY;
                    int bk = 0;
#pragma unroll
                    for (int k = 0; k < NMAX; ++k) {
                        sl[k] = 0.0;
                        if (k < n) {
                            double s = 0.0;
                            #pragma unroll 1
                            for (int j = 0; j <= i; ++j) s = fma(T[(size_t)ind[j] * ldT + k], u[j], s);
                            sl[k] = s;
                            double t = INFINITY;
                            if (!((act >> k) & 1u) && s < 1.0) t = (cc - DtR[k]) / (1.0 - s);
                            if (k == 0 || fabs(t) < best) { best = fabs(t); step = t; bk = k; }
                        }
                    }
                    cur = bk;
                    double coeff1 = 0.0, coeff2 = 0.0;
                    #pragma unroll 1
                    for (int j = 0; j <= i; ++j) {
                        const double dl = DtR[ind[j]];
                        coeff1 += dl > 0.0 ? u[j] : -u[j];
                        coeff2 = fma(dl, u[j], coeff2);
                    }
                    const double step_max2 = cc - lambda1;
                    step = fmin(fmin(step, step_max2), step_max);
                    if (step == INFINITY) break;
                    #pragma unroll 1
                    for (int j = 0; j <= i; ++j) {
                        coef[j] = fma(step, u[j], coef[j]);
                        if (coef[j] < 0.0) coef[j] = 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < NMAX; ++k)
                        if (k < n) DtR[k] = fma(-step, sl[k], DtR[k]);
                    normX += coeff1 * step * step - 2.0 * coeff2 * step;
                    if (step == step_max) {
                        // remove active position fz: shrink ind / coef and downdate the inverse
                        const int z = fz;
                        act &= ~(1u << ind[z]);
                        const double schur_r = Mi[z * LMAX + z];
                        #pragma unroll 1
                        for (int k = 0; k < i; ++k) u[k] = Mi[(k < z ? k : k + 1) * LMAX + z];
                        #pragma unroll 1
                        for (int j = z; j < i; ++j) { ind[j] = ind[j + 1]; coef[j] = coef[j + 1]; }
                        ind[i] = -1; coef[i] = 0.0;
                        #pragma unroll 1
                        for (int r = 0; r < i; ++r) {
                            const int rs = r < z ? r : r + 1;
                            #pragma unroll 1
                            for (int c = 0; c < i; ++c) {
                                const int cs = c < z ? c : c + 1;
                                Mi[r * LMAX + c] = Mi[rs * LMAX + cs] - u[r] * u[c] / schur_r;
                            }
                        }
                        newAtom = 0;
                        na = i;
                        i -= 2;
                    } else {
                        newAtom = 1;
                    }
                    if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == L - 1) break;
                }
                #pragma unroll 1
                for (int j = 0; j < na; ++j)
                    if (ind[j] >= 0) x[ind[j]] = coef[j];
            }
        }
        // ---- maps and optional outputs
        double xr[NMAX];  // coefficients as the fit returned them (SANDI's maps un-normalise x in place)
#pragma unroll
        for (int k = 0; k < NMAX; ++k) xr[k] = x[k];
        const int support = small_maps<MODEL, NMAX>(p, x, vox);
        if (p.support_out) p.support_out[vox] = support;
        if (p.coeff_out) {
#pragma unroll
            for (int k = 0; k < NMAX; ++k)
                if (k < n) p.coeff_out[vox * n + k] = x[k];
        }
        if (p.flags & (FLAG_RMSE | FLAG_NRMSE | FLAG_EXTRA)) {
            // fit errors (amico/models.pyx:45-71; SANDI: normalised A with the rescaled x, reference quirk iv) and FreeWater's corrected DWI
            double den = 0.0, acc = 0.0;
            #pragma unroll 1
            for (int r = 0; r < m; ++r) {
                const double yr = p.y_f64 ? yd[r] : (double)yf[r];
                const TS *row = S + (size_t)r * n_pad;
                double ye = 0.0, fw = 0.0;
#pragma unroll
                for (int k = 0; k < NMAX; ++k)
                    if (k < n) {
                        if (x[k] != 0.0) ye = fma((double)row[k], x[k], ye);
                        if (MODEL == MODEL_FREEWATER && k >= n - p.n_iso) fw = fma((double)row[k], xr[k], fw);
                    }
                const double d = yr - ye;
                acc = fma(d, d, acc);
                den = fma(yr, yr, den);
                if (MODEL == MODEL_FREEWATER && (p.flags & FLAG_EXTRA)) {
                    const double cv = yr - fw;
                    p.extra[vox * m + r] = cv < 0.0 ? 0.0 : cv;
                }
            }
            if (p.flags & FLAG_RMSE) p.rmse[vox] = sqrt(acc / (double)m);
            if (p.flags & FLAG_NRMSE) p.nrmse[vox] = den > 1e-16 ? sqrt(acc / den) : 0.0;
        }
    }
}

}  // namespace amx
