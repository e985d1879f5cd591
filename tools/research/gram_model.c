/* Algorithm model (research tool, not product, not oracle): Gram-space versions of the two solvers
 * as the CUDA kernels run them, in plain C loops.  Used to measure parity rates against the oracle
 * before/while writing the kernels. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAP 64
double gm_thr = 1e-6;   /* A-space re-evaluation threshold on d2 / H_jj (set from Python) */
double gm_floor = 0.0;
long gm_ill_solves = 0, gm_ill_nnls = 0, gm_nnls_calls = 0;  /* refined passive solves / NNLS calls that took a near-dependent atom / all */
int gm_lars_incr = 0;     /* LARS: path direction u = invGs 1 updated in O(|S|) per step instead of recomputed in O(|S|^2) */
long gm_lars_signflips = 0;  /* steps where an active correlation was not positive (the incremental form assumes +1) */  /* refined candidates with d2 <= gm_floor * H_jj count as dependent */

/* ---- NNLS, Lawson-Hanson pivoting in Gram space; Cholesky factor L of H_PP (row-major, CAP ld),
 * appended row-wise on entry, rebuilt from H_PP after removals.  z = L^-1 c_P kept incrementally.
 * mode bit0: CSNE refinement of each passive solve.  bit1: explicit W=L^-1 instead of substitution. */
static void chol_rebuild(const double *H, int ld, const int *P, int np_, double L[CAP][CAP], const double *c, double *z)
{
    int i, j, k;
    for (i = 0; i < np_; ++i) {
        for (j = 0; j <= i; ++j) {
            double s = H[(size_t)P[i] * ld + P[j]];
            for (k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? sqrt(s > 0 ? s : 1e-300) : s / L[j][j];
        }
        { double s = c[P[i]]; for (k = 0; k < i; ++k) s -= L[i][k] * z[k]; z[i] = s / L[i][i]; }
    }
}
static void back_subst(double L[CAP][CAP], int np_, const double *z, double *s)
{
    int i, k; double t[CAP];
    for (i = 0; i < np_; ++i) t[i] = z[i];
    for (k = np_ - 1; k >= 0; --k) { s[k] = t[k] / L[k][k]; for (i = 0; i < k; ++i) t[i] -= L[k][i] * s[k]; }
}
static void fwd_subst(double L[CAP][CAP], int np_, const double *g, double *v)
{
    int i, k; double t[CAP];
    for (i = 0; i < np_; ++i) t[i] = g[i];
    for (k = 0; k < np_; ++k) { v[k] = t[k] / L[k][k]; for (i = k + 1; i < np_; ++i) t[i] -= L[i][k] * v[k]; }
}
int gm_nnls(const double *H, int ld, const double *c, int n, int mcap, double *x,
            const double *A, const double *y, int m, int mode, int *stats)
{
    int P[CAP], inP[512], np_ = 0, iter = 0, itmax = 3 * n, outer = 0, nrem = 0, ill = 0;  /* ill: a near-dependent atom was accepted */
    static __thread double L[CAP][CAP];
    double w[512], s[CAP], g[CAP], v[CAP], z[CAP];
    int i, j, k, a;
    for (j = 0; j < n; ++j) { x[j] = 0; inP[j] = 0; }
    while (np_ < n && np_ < mcap && np_ < CAP) {
        int aspace_dual = 0;
        if ((mode & 16) && np_ > 0) {
            /* exact-fit regime: once the residual has collapsed, c - Hx is rounding noise at 1e-16 |c|; the reference's dual
             * A'(y - Ax) is not.  Residual in A-space (cheap: m x np), switch when |r|^2 < 1e-12 |y|^2. */
            double r2 = 0, y2 = 0;
            static __thread double rr[1024];
            for (i = 0; i < m; ++i) { double t = y[i]; for (k = 0; k < np_; ++k) t -= A[(size_t)P[k] * m + i] * x[P[k]]; rr[i] = t; r2 += t * t; y2 += y[i] * y[i]; }
            if (r2 < 1e-12 * y2) {
                aspace_dual = 1;
                for (j = 0; j < n; ++j) {
                    if (inP[j]) { w[j] = 0; continue; }
                    double acc = 0; for (i = 0; i < m; ++i) acc += A[(size_t)j * m + i] * rr[i];
                    w[j] = acc;
                }
                if (stats) stats[10]++;
            }
        }
        if (!aspace_dual)
        for (j = 0; j < n; ++j) {
            if (inP[j]) { w[j] = 0; continue; }
            double acc = c[j];
            for (k = 0; k < np_; ++k) acc -= H[(size_t)P[k] * ld + j] * x[P[k]];
            w[j] = acc;
        }
        double d2 = 0, znew = 0;
        for (;;) {
            double wmax = 0; j = -1;
            for (k = 0; k < n; ++k) if (!inP[k] && w[k] > wmax) { wmax = w[k]; j = k; }
            if (j < 0) goto done;
            double vv = 0, vz = 0;
            for (a = 0; a < np_; ++a) g[a] = H[(size_t)P[a] * ld + j];
            fwd_subst(L, np_, g, v);
            for (a = 0; a < np_; ++a) { vv += v[a] * v[a]; vz += v[a] * z[a]; }
            d2 = H[(size_t)j * ld + j] - vv;
            if (stats) { double rel = d2 / H[(size_t)j * ld + j]; stats[3]++; if (rel < 1e-6) stats[4]++; if (rel < 1e-8) stats[5]++; if (rel < 1e-10) stats[6]++; if (rel<1e-12) stats[7]++; }
            double znum = c[j] - vz; int ill_cand = 0;
            if ((mode & 8) && np_ > 0 && d2 < gm_thr * H[(size_t)j * ld + j]) {
                /* A-space re-evaluation of a near-dependent candidate: r = a_j - A_P beta, beta = L^-T v;
                 * d2 = |r|^2 and the numerator r.y come without the cancellation of H_jj - v.v */
                double beta[CAP], rr, acc2 = 0, accy = 0;
                back_subst(L, np_, v, beta);
                for (i = 0; i < m; ++i) {
                    rr = A[(size_t)j * m + i];
                    for (a = 0; a < np_; ++a) rr -= A[(size_t)P[a] * m + i] * beta[a];
                    acc2 += rr * rr; accy += rr * y[i];
                }
                d2 = acc2; znum = accy; ill_cand = 1;
                if (d2 <= gm_floor * H[(size_t)j * ld + j]) d2 = 0;
                if (stats) stats[9]++;
            }
            if (d2 > 0) {
                double unorm = sqrt(vv), t = unorm + sqrt(d2) * 0.01;
                if (t - unorm > 0) {
                    znew = znum / sqrt(d2);
                    if (znew > 0) { if (ill_cand) ill = 1; break; }   /* ztest = znew / sqrt(d2) */
                }
            }
            w[j] = 0;
        }
        for (a = 0; a < np_; ++a) L[np_][a] = v[a];
        L[np_][np_] = sqrt(d2); z[np_] = znew;
        P[np_++] = j; inP[j] = 1; ++outer;
        if (stats && np_ > stats[8]) stats[8] = np_;
        for (;;) {
            if (++iter > itmax) goto done;
            back_subst(L, np_, z, s);
            if ((mode & 32) && ill) gm_ill_solves++;
            if ((mode & 1) || ((mode & 32) && ill)) {
                double r[1024], q[CAP], dz[CAP], ds[CAP];
                for (i = 0; i < m; ++i) { double t = y[i]; for (a = 0; a < np_; ++a) t -= A[(size_t)P[a] * m + i] * s[a]; r[i] = t; }
                for (a = 0; a < np_; ++a) { double t = 0; for (i = 0; i < m; ++i) t += A[(size_t)P[a] * m + i] * r[i]; q[a] = t; }
                fwd_subst(L, np_, q, dz); back_subst(L, np_, dz, ds);
                for (a = 0; a < np_; ++a) s[a] += ds[a];
            }
            double alpha = 2.0; int jj = -1;
            for (a = 0; a < np_; ++a) if (s[a] <= 0) {
                double t = -x[P[a]] / (s[a] - x[P[a]]);
                if (alpha > t) { alpha = t; jj = a; }
            }
            if (jj < 0) break;
            for (a = 0; a < np_; ++a) x[P[a]] += alpha * (s[a] - x[P[a]]);
            x[P[jj]] = 0;
            int removed[CAP], nremoved = 0, oldnp = np_;
            { int q = 0; for (a = 0; a < np_; ++a) { if (x[P[a]] <= 0) { inP[P[a]] = 0; x[P[a]] = 0; ++nrem; removed[nremoved++] = a; } else P[q++] = P[a]; } np_ = q; }
            if (np_ == 0) break;
            if (mode & 4) {
                /* Givens downdate: rem[] holds removed positions (ascending) of the OLD ordering */
                for (int ri = nremoved - 1; ri >= 0; --ri) {
                    int q = removed[ri], pn = oldnp;  /* current size before this removal */
                    /* delete row q */
                    for (int r2 = q; r2 < pn - 1; ++r2) { for (int c2 = 0; c2 <= r2 + 1; ++c2) L[r2][c2] = L[r2 + 1][c2]; }
                    for (int r2 = q; r2 < pn - 1; ++r2) {
                        double a_ = L[r2][r2], b_ = L[r2][r2 + 1], rho = sqrt(a_ * a_ + b_ * b_), cs = a_ / rho, sn = b_ / rho;
                        for (int i2 = r2; i2 < pn - 1; ++i2) {
                            double u1 = L[i2][r2], u2 = L[i2][r2 + 1];
                            L[i2][r2] = cs * u1 + sn * u2; L[i2][r2 + 1] = -sn * u1 + cs * u2;
                        }
                        double z1 = z[r2], z2 = z[r2 + 1];
                        z[r2] = cs * z1 + sn * z2; z[r2 + 1] = -sn * z1 + cs * z2;
                    }
                    oldnp = pn - 1;
                }
            } else
            chol_rebuild(H, ld, P, np_, L, c, z);
        }
        for (a = 0; a < np_; ++a) x[P[a]] = s[a];
    }
done:
    gm_nnls_calls++; if (ill) gm_ill_nnls++;
    if (stats) { stats[0] += outer; stats[1] += iter; stats[2] += nrem; }
    return outer;
}

/* ---- LARS in Gram space: same path as oracle lars_core, Gram columns from a table ------------ */
int gm_lars(const double *G, int ld, double ridge_in, const double *DtR_in, double normX, int K, int L,
            double lambda1, double *x)
{
    int LL, length_path, i, j, k, iter = 0, currentInd, newAtom = 1, first_zero, index;
    double ridge = ridge_in > 1e-10 ? ridge_in : 1e-10;
    double step, step_max, step_max2, cc, coeff1, coeff2, best;
    double DtR[512], Gs[CAP * CAP], invGs[CAP * CAP], u[512], work[3 * 512], coeffs[CAP];
    const double *Ga[CAP];
    int ind[CAP];
    double rs[CAP];  /* incremental row sums of invGs */
    if (L > K) L = K;
    if (L > CAP) L = CAP;
    LL = L; length_path = 4 * L;
    for (k = 0; k < K; ++k) { x[k] = 0; DtR[k] = DtR_in[k]; }
    if (L <= 0) return 0;
    for (j = 0; j < L; ++j) { coeffs[j] = 0; ind[j] = -1; }
    currentInd = 0;
    for (k = 1; k < K; ++k) if (DtR[k] > DtR[currentInd]) currentInd = k;
    if (fabs(DtR[currentInd]) < lambda1) return 0;
#define GA(jj, kk) (Ga[jj][kk] + ((kk) == ind[jj] ? ridge : 0.0))
    for (i = 0; i < L; ++i) {
        ++iter;
        if (newAtom) {
            ind[i] = currentInd;
            Ga[i] = G + (size_t)currentInd * ld;
            for (j = 0; j <= i; ++j) Gs[i * LL + j] = GA(i, ind[j]);
            if (i == 0) invGs[0] = 1.0 / Gs[0];
            else {
                double schur, dot = 0;
                for (int r = 0; r < i; ++r) { double s = 0; for (int c = 0; c < i; ++c) s += (r <= c ? invGs[c * LL + r] : invGs[r * LL + c]) * Gs[i * LL + c]; u[r] = s; }
                for (j = 0; j < i; ++j) dot += u[j] * Gs[i * LL + j];
                schur = 1.0 / (Gs[i * LL + i] - dot);
                invGs[i * LL + i] = schur;
                for (j = 0; j < i; ++j) invGs[i * LL + j] = -schur * u[j];
                for (k = 0; k < i; ++k) for (j = 0; j <= k; ++j) invGs[k * LL + j] += schur * u[j] * u[k];
                { double U = 0; for (j = 0; j < i; ++j) U += u[j];
                  for (j = 0; j < i; ++j) rs[j] += schur * u[j] * (U - 1.0);
                  rs[i] = schur * (1.0 - U); }
            }
            if (i == 0) rs[0] = invGs[0];
        }
        for (j = 0; j <= i; ++j) work[j] = DtR[ind[j]] > 0 ? 1.0 : -1.0;
        for (int r = 0; r <= i; ++r) { double s = 0; for (int c = 0; c <= i; ++c) s += (r <= c ? invGs[c * LL + r] : invGs[r * LL + c]) * work[c]; u[r] = s; }
        { int allpos = 1; for (j = 0; j <= i; ++j) if (work[j] < 0) allpos = 0;
          if (!allpos) gm_lars_signflips++;
          if (gm_lars_incr && allpos) for (j = 0; j <= i; ++j) u[j] = rs[j]; }
        step_max = INFINITY; first_zero = -1;
        for (j = 0; j <= i; ++j) { double ratio = -coeffs[j] / u[j]; if (ratio > 0 && ratio <= step_max) { step_max = ratio; first_zero = j; } }
        cc = fabs(DtR[ind[0]]);
        for (k = 0; k < K; ++k) { double s = 0; for (j = 0; j <= i; ++j) s += GA(j, k) * u[j]; work[2 * K + k] = s; }
        for (k = 0; k < K; ++k) work[K + k] = work[2 * K + k];
        for (j = 0; j <= i; ++j) work[K + ind[j]] = INFINITY;
        for (k = 0; k < K; ++k) work[K + k] = (work[K + k] < INFINITY && work[K + k] < 1.0) ? (cc - DtR[k]) / (1.0 - work[K + k]) : INFINITY;
        index = 0; best = fabs(work[K]);
        for (k = 1; k < K; ++k) if (fabs(work[K + k]) < best) { best = fabs(work[K + k]); index = k; }
        step = work[K + index];
        currentInd = index;
        coeff1 = 0; coeff2 = 0;
        for (j = 0; j <= i; ++j) coeff1 += DtR[ind[j]] > 0 ? u[j] : -u[j];
        for (j = 0; j <= i; ++j) coeff2 += DtR[ind[j]] * u[j];
        step_max2 = cc - lambda1;
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        for (j = 0; j <= i; ++j) coeffs[j] += step * u[j];
        for (j = 0; j <= i; ++j) if (coeffs[j] < 0) coeffs[j] = 0;
        for (k = 0; k < K; ++k) DtR[k] -= step * work[2 * K + k];
        normX += coeff1 * step * step - 2 * coeff2 * step;
        if (step == step_max) {
            int z = first_zero; double schur;
            for (j = z; j < i; ++j) { Ga[j] = Ga[j + 1]; ind[j] = ind[j + 1]; coeffs[j] = coeffs[j + 1]; }
            ind[i] = -1; coeffs[i] = 0;
            for (j = z; j < i; ++j) {
                for (k = 0; k < z; ++k) Gs[j * LL + k] = Gs[(j + 1) * LL + k];
                for (k = z; k < i; ++k) Gs[j * LL + k] = Gs[(j + 1) * LL + k + 1];
            }
            schur = invGs[z * LL + z];
            for (k = 0; k < z; ++k) u[k] = invGs[z * LL + k];
            for (k = z; k < i; ++k) u[k] = invGs[(k + 1) * LL + z];
            for (j = z; j < i; ++j) {
                for (k = 0; k < z; ++k) invGs[j * LL + k] = invGs[(j + 1) * LL + k];
                for (k = z; k < i; ++k) invGs[j * LL + k] = invGs[(j + 1) * LL + k + 1];
            }
            for (k = 0; k < i; ++k) for (j = 0; j <= k; ++j) invGs[k * LL + j] -= u[j] * u[k] / schur;
            { /* row sums: drop row/column z, then the rank-1 downdate */
              double S = 0; for (k = 0; k < i; ++k) S += u[k];
              for (k = z; k < i; ++k) rs[k] = rs[k + 1];
              for (k = 0; k < i; ++k) rs[k] = rs[k] - u[k] - u[k] * S / schur; }
            newAtom = 0; i -= 2;
        } else newAtom = 1;
        if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == L - 1) break;
    }
    for (j = 0; j < L; ++j) if (ind[j] >= 0) x[ind[j]] = coeffs[j];
    return iter;
}

/* ---- NODDI pipeline on precomputed per-direction tables --------------------------------------
 * A: m x n col-major double for this voxel's direction; T1 = A'A (n x n); T2 = A2'A2 (nwm x nwm). */
void gm_noddi_voxel(const double *A, const double *T1, const double *T2, const double *y, int m, int n_wm, int exvivo,
                    const int64_t *dwi_idx, int dc, const double *norms /* dc x n_wm */, const double *iso,
                    double l1, double l2, int mode, double *x_out, int *sup_out)
{
    int n = n_wm + 1 + (exvivo ? 1 : 0), j, k, r;
    double c1[512], c2[512], y2[1024], x[512], x3[512], H3[64 * 64], c3[64], A3[64 * 1024 > 1 ? 1 : 1];
    int pos[512], pc = 0;
    (void)A3;
    for (k = 0; k < n; ++k) { double s = 0; for (r = 0; r < m; ++r) s += A[(size_t)k * m + r] * y[r]; c1[k] = s; }
    gm_nnls(T1, n, c1, n, m, x, A, y, m, mode, NULL);
    double xiso = x[n - 1], xdot = exvivo ? x[n - 2] : 0, nx = 0;
    for (j = 0; j < dc; ++j) {
        r = (m == 1 + dc) ? j + 1 : (int)dwi_idx[j];
        y2[j] = y[r] - xiso * iso[r];
        if (exvivo) y2[j] -= xdot;
        if (y2[j] < 0) y2[j] = 0;
        nx += y2[j] * y2[j];
    }
    for (k = 0; k < n_wm; ++k) {
        double s = 0;
        for (j = 0; j < dc; ++j) { r = (m == 1 + dc) ? j + 1 : (int)dwi_idx[j]; s += (A[(size_t)k * m + r] * norms[(size_t)j * n_wm + k]) * y2[j]; }
        c2[k] = s;
    }
    gm_lars(T2, n_wm, l2, c2, nx, n_wm, dc < n_wm ? dc : n_wm, l1, x);
    if (exvivo) x[n - 2] = 1; x[n - 1] = 1;
    for (j = 0; j < n; ++j) if (x[j] > 0) pos[pc++] = j;
    if (sup_out) *sup_out = pc;
    {
        double *Asub = (double *)malloc(sizeof(double) * (size_t)m * pc);
        for (j = 0; j < pc; ++j) {
            c3[j] = c1[pos[j]];
            for (k = 0; k < pc; ++k) H3[j * pc + k] = T1[(size_t)pos[j] * n + pos[k]];
            memcpy(Asub + (size_t)j * m, A + (size_t)pos[j] * m, sizeof(double) * m);
        }
        gm_nnls(H3, pc, c3, pc, m, x3, Asub, y, m, mode, NULL);
        free(Asub);
    }
    for (j = 0; j < pc; ++j) x[pos[j]] = x3[j];
    for (j = 0; j < n; ++j) x_out[j] = x[j];
}
