/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it, and only as the checker / the timed CPU baseline.
 *
 * Plain-C restatement of the reference's per-voxel fit (daducci/AMICO v2.1.0):
 *   - orc_dir_to_lut_idx      <- amico/lut.pyx:314-356
 *   - orc_fit_noddi           <- amico/models.pyx:816-991   (loop 901-981)
 *   - orc_fit_freewater       <- amico/models.pyx:1168-1286 (loop 1231-1276)
 *   - orc_fit_sandi           <- amico/models.pyx:1509-1627 (loop 1567-1619)
 *   - orc_fit_czb             <- amico/models.pyx:546-652   (loop 607-644)
 *   - rmse / nrmse            <- amico/models.pyx:45-71
 *   - chunking over threads   <- amico/models.pyx:204-211
 *
 * The two solvers the reference calls (`from cyspams.interfaces cimport nnls, lasso`,
 * amico/models.pyx:18; call sites :615, :911, :926, :940, :1238, :1569) live in the third-party
 * package spams-cython (PyPI, ">=1.0.0", un-pinned: pyproject.toml:5; source NOT under
 * /root/reference).  They are restated here from the published algorithms:
 *   - orc_nnls  : Lawson & Hanson, "Solving Least Squares Problems" (1974), ch. 23, algorithm NNLS
 *                 (Householder QR on the passive set, entering index = argmax of the dual, linear
 *                 independence + positivity test of the candidate, Givens removal), max 3n iterations.
 *   - orc_lasso : SPAMS `lasso` (Mairal et al.), mode=PENALTY, pos=true: LARS/homotopy on the Gram
 *                 matrix G = A'A + max(lambda2,1e-10) I with explicit (A_S'A_S)^-1 updates
 *                 ("coreLARS2"), path truncated at L = min(m, n) active atoms and 4L steps.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer fixtures for this
 * path and spams-cython cannot be installed here, so the solver
 * restatements are checked only against independent implementations (scipy.optimize.nnls, KKT
 * conditions) and the model glue against the reference's own Cython compiled with these solvers
 * (oracle/_ref, see oracle/build_ref.py).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ dir -> LUT index */
/* amico/lut.pyx:314-356.  Flips `d` in place when d[1] < 0 (the reference does the same on its
 * view of DIRs).  Returns the LUT index, or -1 when the whole-degree angles leave [0,180]
 * (the reference raises RuntimeError there); ii1_out / ii2_out receive the angles when non-NULL. */
int orc_dir_to_lut_idx(double *d, const int16_t *htable, int *ii1_out, int *ii2_out)
{
    double i1, i2;
    if (d[1] < 0.0) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; }
    i2 = fmod(atan2(d[1], d[0]), 2.0 * M_PI);
    if (i2 < 0.0) i2 = fmod(i2 + 2.0 * M_PI, 2.0 * M_PI);
    if (i2 > M_PI) {
        i2 = fmod(atan2(-d[1], -d[0]), 2.0 * M_PI);
        i1 = atan2(sqrt(d[0] * d[0] + d[1] * d[1]), -d[2]);
    } else {
        i1 = atan2(sqrt(d[0] * d[0] + d[1] * d[1]), d[2]);
    }
    {
        double r1 = round(i1 / M_PI * 180.0), r2 = round(i2 / M_PI * 180.0);
        int a, b;
        /* NaN or out-of-range angles: the C cast in the reference is undefined for NaN; it then
         * fails the range test on every platform we know of.  Map both to the error return. */
        if (!(r1 >= 0.0 && r1 <= 180.0 && r2 >= 0.0 && r2 <= 180.0)) {
            if (ii1_out) *ii1_out = (r1 == r1) ? (int)fmax(fmin(r1, 1e9), -1e9) : -1;
            if (ii2_out) *ii2_out = (r2 == r2) ? (int)fmax(fmin(r2, 1e9), -1e9) : -1;
            return -1;
        }
        a = (int)r1; b = (int)r2;
        if (ii1_out) *ii1_out = a;
        if (ii2_out) *ii2_out = b;
        return (int)htable[a * 181 + b];
    }
}

/* ------------------------------------------------------------------ Lawson-Hanson NNLS */
/* Householder "H12": construct (mode 1) the reflector that zeroes u[l1..m) against pivot lp, or
 * apply (mode 2) a previously constructed one to the vector c. */
static void h12_construct(int lp, int l1, int m, double *u, double *up)
{
    double cl = fabs(u[lp]), sm, clinv;
    int j;
    if (l1 >= m) { /* nothing below the pivot: identity (book: returns when l1 > m) */
        *up = 0.0;
        return;
    }
    for (j = l1; j < m; ++j) if (fabs(u[j]) > cl) cl = fabs(u[j]);
    if (cl <= 0.0) { *up = 0.0; return; }
    clinv = 1.0 / cl;
    sm = (u[lp] * clinv) * (u[lp] * clinv);
    for (j = l1; j < m; ++j) sm += (u[j] * clinv) * (u[j] * clinv);
    cl *= sqrt(sm);
    if (u[lp] > 0.0) cl = -cl;
    *up = u[lp] - cl;
    u[lp] = cl;
}

static void h12_apply(int lp, int l1, int m, const double *u, double up, double *c)
{
    double b = up * u[lp], sm;
    int i;
    if (l1 >= m) return;
    if (b >= 0.0) return;
    b = 1.0 / b;
    sm = c[lp] * up;
    for (i = l1; i < m; ++i) sm += c[i] * u[i];
    if (sm != 0.0) {
        sm *= b;
        c[lp] += sm * up;
        for (i = l1; i < m; ++i) c[i] += sm * u[i];
    }
}

static void givens(double a, double b, double *c, double *s, double *sig)
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

typedef struct {
    double *a, *b, *w, *zz;
    int *index;
    int cap_m, cap_n;
} nnls_ws;

static void nnls_ws_init(nnls_ws *ws) { memset(ws, 0, sizeof(*ws)); }
static void nnls_ws_free(nnls_ws *ws)
{
    free(ws->a); free(ws->b); free(ws->w); free(ws->zz); free(ws->index);
    memset(ws, 0, sizeof(*ws));
}
static void nnls_ws_reserve(nnls_ws *ws, int m, int n)
{
    if (m <= ws->cap_m && n <= ws->cap_n) return;
    if (m < ws->cap_m) m = ws->cap_m;
    if (n < ws->cap_n) n = ws->cap_n;
    nnls_ws_free(ws);
    ws->a = (double *)malloc(sizeof(double) * (size_t)m * n);
    ws->b = (double *)malloc(sizeof(double) * m);
    ws->zz = (double *)malloc(sizeof(double) * m);
    ws->w = (double *)malloc(sizeof(double) * n);
    ws->index = (int *)malloc(sizeof(int) * n);
    ws->cap_m = m; ws->cap_n = n;
}

/* min ||A x - y||_2  s.t. x >= 0.  A is m x n column-major and, like y, left untouched (the
 * reference reuses both after the call: amico/models.pyx:917-921, 939-940, 971).
 * Returns the number of outer iterations (entering indices); <0 if the 3n iteration cap hit. */
static int nnls_core(nnls_ws *ws, const double *A_in, const double *y_in, int m, int n, double *x, double *rnorm)
{
    double *a, *b, *w, *zz, up = 0.0, unorm, asave, ztest, alpha, t, cc, ss, sig, tmp;
    int *index, iz1 = 0, iz2 = n - 1, nsetp = 0, npp1 = 0, iter = 0, itmax = 3 * n, outer = 0;
    int i, j = 0, l, iz, izmax = 0, jz, jj = 0, ip, ii, ok = 1;
    nnls_ws_reserve(ws, m, n);
    a = ws->a; b = ws->b; w = ws->w; zz = ws->zz; index = ws->index;
    memcpy(a, A_in, sizeof(double) * (size_t)m * n);
    memcpy(b, y_in, sizeof(double) * m);
#define AE(r, c) a[(size_t)(c) * m + (r)]
    for (i = 0; i < n; ++i) { x[i] = 0.0; index[i] = i; }

    while (iz1 <= iz2 && nsetp < m) {
        /* dual vector on the zero set */
        for (iz = iz1; iz <= iz2; ++iz) {
            double sm = 0.0;
            j = index[iz];
            for (l = npp1; l < m; ++l) sm += AE(l, j) * b[l];
            w[j] = sm;
        }
        for (;;) {
            double wmax = 0.0;
            for (iz = iz1; iz <= iz2; ++iz) {
                j = index[iz];
                if (w[j] > wmax) { wmax = w[j]; izmax = iz; }
            }
            if (wmax <= 0.0) goto done;
            iz = izmax; j = index[iz];
            /* candidate column: independence test, then sign of its would-be coefficient */
            asave = AE(npp1, j);
            h12_construct(npp1, npp1 + 1, m, &AE(0, j), &up);
            unorm = 0.0;
            for (l = 0; l < nsetp; ++l) unorm += AE(l, j) * AE(l, j);
            unorm = sqrt(unorm);
            tmp = unorm + fabs(AE(npp1, j)) * 0.01;
            if (tmp - unorm > 0.0) {
                memcpy(zz, b, sizeof(double) * m);
                h12_apply(npp1, npp1 + 1, m, &AE(0, j), up, zz);
                ztest = zz[npp1] / AE(npp1, j);
                if (ztest > 0.0) break;
            }
            AE(npp1, j) = asave;
            w[j] = 0.0;
        }
        /* move j from the zero set to the passive set */
        memcpy(b, zz, sizeof(double) * m);
        index[iz] = index[iz1]; index[iz1] = j;
        ++iz1; nsetp = npp1 + 1; ++npp1; ++outer;
        for (jz = iz1; jz <= iz2; ++jz) {
            jj = index[jz];
            h12_apply(nsetp - 1, npp1, m, &AE(0, j), up, &AE(0, jj));
        }
        for (l = npp1; l < m; ++l) AE(l, j) = 0.0;
        w[j] = 0.0;
        /* triangular solve into zz */
        memcpy(zz, b, sizeof(double) * m);
        for (ip = nsetp - 1; ip >= 0; --ip) {
            if (ip != nsetp - 1) for (ii = 0; ii <= ip; ++ii) zz[ii] -= AE(ii, jj) * zz[ip + 1];
            jj = index[ip];
            zz[ip] /= AE(ip, jj);
        }
        /* secondary loop */
        for (;;) {
            if (++iter > itmax) { ok = 0; goto done; }
            alpha = 2.0; jj = -1;
            for (ip = 0; ip < nsetp; ++ip) {
                l = index[ip];
                if (zz[ip] <= 0.0) {
                    t = -x[l] / (zz[ip] - x[l]);
                    if (alpha > t) { alpha = t; jj = ip; }
                }
            }
            if (alpha == 2.0) break;
            for (ip = 0; ip < nsetp; ++ip) { l = index[ip]; x[l] += alpha * (zz[ip] - x[l]); }
            /* move index[jj] from passive to zero set, restoring triangularity with Givens */
            i = index[jj];
            for (;;) {
                x[i] = 0.0;
                if (jj != nsetp - 1) {
                    ++jj;
                    for (j = jj; j < nsetp; ++j) {
                        ii = index[j]; index[j - 1] = ii;
                        givens(AE(j - 1, ii), AE(j, ii), &cc, &ss, &sig);
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
                npp1 = nsetp - 1; --nsetp; --iz1; index[iz1] = i;
                /* every remaining passive coefficient should be feasible; if round-off left
                 * one non-positive, move it out too */
                for (jj = 0; jj < nsetp; ++jj) { i = index[jj]; if (x[i] <= 0.0) break; }
                if (jj == nsetp) break;
            }
            memcpy(zz, b, sizeof(double) * m);
            for (ip = nsetp - 1; ip >= 0; --ip) {
                if (ip != nsetp - 1) for (ii = 0; ii <= ip; ++ii) zz[ii] -= AE(ii, jj) * zz[ip + 1];
                jj = index[ip];
                zz[ip] /= AE(ip, jj);
            }
        }
        for (ip = 0; ip < nsetp; ++ip) x[index[ip]] = zz[ip];
    }
done:
    {
        double sm = 0.0;
        for (i = npp1; i < m; ++i) sm += b[i] * b[i];
        if (rnorm) *rnorm = sqrt(sm);
    }
#undef AE
    return ok ? outer : -outer - 1;
}

int orc_nnls(const double *A, const double *y, int m, int n, double *x, double *rnorm)
{
    nnls_ws ws; int r;
    nnls_ws_init(&ws);
    r = nnls_core(&ws, A, y, m, n, x, rnorm);
    nnls_ws_free(&ws);
    return r;
}

/* ------------------------------------------------------------------ SPAMS lasso (LARS, PENALTY, pos) */
typedef struct {
    double *DtR, *Ga, *Gs, *invGs, *u, *work, *coeffs, *sgn;
    int *ind;
    int cap_K, cap_L;
} lars_ws;

static void lars_ws_init(lars_ws *ws) { memset(ws, 0, sizeof(*ws)); }
static void lars_ws_free(lars_ws *ws)
{
    free(ws->DtR); free(ws->Ga); free(ws->Gs); free(ws->invGs); free(ws->u); free(ws->work);
    free(ws->coeffs); free(ws->sgn); free(ws->ind);
    memset(ws, 0, sizeof(*ws));
}
static void lars_ws_reserve(lars_ws *ws, int K, int L)
{
    if (K <= ws->cap_K && L <= ws->cap_L) return;
    if (K < ws->cap_K) K = ws->cap_K;
    if (L < ws->cap_L) L = ws->cap_L;
    lars_ws_free(ws);
    ws->DtR = (double *)malloc(sizeof(double) * K);
    ws->Ga = (double *)malloc(sizeof(double) * (size_t)K * L);
    ws->Gs = (double *)malloc(sizeof(double) * (size_t)L * L);
    ws->invGs = (double *)malloc(sizeof(double) * (size_t)L * L);
    ws->u = (double *)malloc(sizeof(double) * (L > K ? L : K));
    ws->work = (double *)malloc(sizeof(double) * (size_t)K * 3);
    ws->coeffs = (double *)malloc(sizeof(double) * L);
    ws->sgn = (double *)malloc(sizeof(double) * L);
    ws->ind = (int *)malloc(sizeof(int) * L);
    ws->cap_K = K; ws->cap_L = L;
}

/* One Gram column: G[:, j] = A' A[:, j] + ridge * e_j (SPAMS builds G = D'D once per call and
 * adds max(lambda2,1e-10) on the diagonal; only the columns of entering atoms are ever read). */
static void gram_column(const double *A, int m, int K, int j, double ridge, double *out)
{
    const double *aj = A + (size_t)j * m;
    int k, r;
    for (k = 0; k < K; ++k) {
        const double *ak = A + (size_t)k * m;
        double s = 0.0;
        for (r = 0; r < m; ++r) s += ak[r] * aj[r];
        out[k] = s;
    }
    out[j] += ridge;
}

/* Symmetric (upper-stored, leading dimension ld) matrix-vector product, order n. */
static void symv_upper(const double *S, int ld, int n, const double *v, double *out)
{
    int r, c;
    for (r = 0; r < n; ++r) {
        double s = 0.0;
        for (c = 0; c < n; ++c) s += (r <= c ? S[(size_t)c * ld + r] : S[(size_t)r * ld + c]) * v[c];
        out[r] = s;
    }
}

/* x = argmin 1/2||y - A x||^2 + lambda1 |x|_1 + 1/2 lambda2 ||x||^2, x >= 0, followed along the
 * LARS path exactly as SPAMS does (so early truncation at L atoms / 4L steps is reproduced).
 * `L_cap` <= 0 selects the SPAMS default min(m, K).  Returns the number of path steps taken. */
static int lars_core(lars_ws *ws, const double *A, const double *y, int m, int K, double *x,
                     double lambda1, double lambda2, int L_cap)
{
    int L = L_cap > 0 ? L_cap : (m < K ? m : K);
    int LL, length_path, i, j, k, iter = 0, currentInd, newAtom = 1, first_zero, index;
    double ridge = lambda2 > 1e-10 ? lambda2 : 1e-10;
    double normX = 0.0, thrs = 0.0, step, step_max, step_max2, cc, coeff1, coeff2, best;
    double *DtR, *Ga, *Gs, *invGs, *u, *work, *coeffs;
    int *ind;
    if (L > K) L = K;
    LL = L;
    length_path = 4 * L;
    lars_ws_reserve(ws, K, L);
    DtR = ws->DtR; Ga = ws->Ga; Gs = ws->Gs; invGs = ws->invGs; u = ws->u; work = ws->work;
    coeffs = ws->coeffs; ind = ws->ind;
    for (k = 0; k < K; ++k) x[k] = 0.0;
    if (L <= 0) return 0;
    for (k = 0; k < m; ++k) normX += y[k] * y[k];
    for (k = 0; k < K; ++k) {
        const double *ak = A + (size_t)k * m;
        double s = 0.0;
        for (j = 0; j < m; ++j) s += ak[j] * y[j];
        DtR[k] = s;
    }
    for (j = 0; j < L; ++j) { coeffs[j] = 0.0; ind[j] = -1; }
    currentInd = 0;
    for (k = 1; k < K; ++k) if (DtR[k] > DtR[currentInd]) currentInd = k;
    if (fabs(DtR[currentInd]) < lambda1) return 0;

    for (i = 0; i < L; ++i) {
        ++iter;
        if (newAtom) {
            ind[i] = currentInd;
            gram_column(A, m, K, currentInd, ridge, Ga + (size_t)i * K);
            for (j = 0; j <= i; ++j) Gs[(size_t)i * LL + j] = Ga[(size_t)i * K + ind[j]];
            if (i == 0) {
                invGs[0] = 1.0 / Gs[0];
            } else {
                double schur, dot = 0.0;
                symv_upper(invGs, LL, i, Gs + (size_t)i * LL, u);
                for (j = 0; j < i; ++j) dot += u[j] * Gs[(size_t)i * LL + j];
                schur = 1.0 / (Gs[(size_t)i * LL + i] - dot);
                invGs[(size_t)i * LL + i] = schur;
                for (j = 0; j < i; ++j) invGs[(size_t)i * LL + j] = -schur * u[j];
                for (k = 0; k < i; ++k)
                    for (j = 0; j <= k; ++j) invGs[(size_t)k * LL + j] += schur * u[j] * u[k];
            }
        }
        /* path direction */
        for (j = 0; j <= i; ++j) work[j] = DtR[ind[j]] > 0 ? 1.0 : -1.0;
        symv_upper(invGs, LL, i + 1, work, u);
        /* largest step before an active coefficient crosses zero */
        step_max = INFINITY; first_zero = -1;
        for (j = 0; j <= i; ++j) {
            double ratio = -coeffs[j] / u[j];
            if (ratio > 0 && ratio <= step_max) { step_max = ratio; first_zero = j; }
        }
        cc = fabs(DtR[ind[0]]);
        /* correlations' slope  Ga u  (kept in work[2K..3K)) */
        for (k = 0; k < K; ++k) {
            double s = 0.0;
            for (j = 0; j <= i; ++j) s += Ga[(size_t)j * K + k] * u[j];
            work[2 * K + k] = s;
        }
        /* step until an inactive atom reaches the common correlation (positive side only) */
        for (k = 0; k < K; ++k) work[K + k] = work[2 * K + k];
        for (j = 0; j <= i; ++j) work[K + ind[j]] = INFINITY;
        for (k = 0; k < K; ++k)
            work[K + k] = (work[K + k] < INFINITY && work[K + k] < 1.0)
                              ? (cc - DtR[k]) / (1.0 - work[K + k]) : INFINITY;
        /* SPAMS takes the entry of smallest magnitude ("iamin"), lowest index on ties */
        index = 0; best = fabs(work[K]);
        for (k = 1; k < K; ++k) if (fabs(work[K + k]) < best) { best = fabs(work[K + k]); index = k; }
        step = work[K + index];
        currentInd = index;
        coeff1 = 0.0; coeff2 = 0.0;
        for (j = 0; j <= i; ++j) coeff1 += DtR[ind[j]] > 0 ? u[j] : -u[j];
        for (j = 0; j <= i; ++j) coeff2 += DtR[ind[j]] * u[j];
        step_max2 = cc - lambda1;
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        for (j = 0; j <= i; ++j) coeffs[j] += step * u[j];
        for (j = 0; j <= i; ++j) if (coeffs[j] < 0) coeffs[j] = 0;
        for (k = 0; k < K; ++k) DtR[k] -= step * work[2 * K + k];
        normX += coeff1 * step * step - 2 * coeff2 * step;
        thrs += step * coeff1;
        if (step == step_max) {
            /* remove atom `first_zero`: shrink Ga, ind, coeffs, Gs, and downdate invGs */
            int z = first_zero;
            double schur;
            for (j = z; j < i; ++j) {
                memcpy(Ga + (size_t)j * K, Ga + (size_t)(j + 1) * K, sizeof(double) * K);
                ind[j] = ind[j + 1];
                coeffs[j] = coeffs[j + 1];
            }
            ind[i] = -1; coeffs[i] = 0;
            for (j = z; j < i; ++j) {
                for (k = 0; k < z; ++k) Gs[(size_t)j * LL + k] = Gs[(size_t)(j + 1) * LL + k];
                for (k = z; k < i; ++k) Gs[(size_t)j * LL + k] = Gs[(size_t)(j + 1) * LL + k + 1];
            }
            schur = invGs[(size_t)z * LL + z];
            for (k = 0; k < z; ++k) u[k] = invGs[(size_t)z * LL + k];
            for (k = z; k < i; ++k) u[k] = invGs[(size_t)(k + 1) * LL + z];
            for (j = z; j < i; ++j) {
                for (k = 0; k < z; ++k) invGs[(size_t)j * LL + k] = invGs[(size_t)(j + 1) * LL + k];
                for (k = z; k < i; ++k) invGs[(size_t)j * LL + k] = invGs[(size_t)(j + 1) * LL + k + 1];
            }
            for (k = 0; k < i; ++k)
                for (j = 0; j <= k; ++j) invGs[(size_t)k * LL + j] -= u[j] * u[k] / schur;
            newAtom = 0;
            i -= 2;
        } else {
            newAtom = 1;
        }
        if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 ||
            i == L - 1)
            break;
    }
    (void)thrs;
    for (j = 0; j < L; ++j) if (ind[j] >= 0) x[ind[j]] = coeffs[j];
    return iter;
}

int orc_lasso(const double *A, const double *y, int m, int n, int p, double *x, double lambda1, double lambda2)
{
    lars_ws ws; int r = 0, s;
    lars_ws_init(&ws);
    for (s = 0; s < p; ++s) r = lars_core(&ws, A, y + (size_t)s * m, m, n, x + (size_t)s * n, lambda1, lambda2, 0);
    lars_ws_free(&ws);
    return r;
}

/* ------------------------------------------------------------------ fit errors (models.pyx:45-71) */
static double fit_rmse(const double *A, int m, int n, const double *y, const double *x, double *yest)
{
    double acc = 0.0; int i, j;
    for (i = 0; i < m; ++i) {
        yest[i] = 0.0;
        for (j = 0; j < n; ++j) yest[i] += A[(size_t)j * m + i] * x[j];
        acc += pow(y[i] - yest[i], 2.0) / m;
    }
    return sqrt(acc);
}

static double fit_nrmse(const double *A, int m, int n, const double *y, const double *x, double *yest)
{
    double den = 0.0, acc = 0.0; int i, j;
    for (i = 0; i < m; ++i) {
        yest[i] = 0.0;
        den += pow(y[i], 2.0);
        for (j = 0; j < n; ++j) yest[i] += A[(size_t)j * m + i] * x[j];
    }
    if (den > 1e-16) {
        for (i = 0; i < m; ++i) acc += pow(y[i] - yest[i], 2.0) / den;
        return sqrt(acc);
    }
    return 0.0;
}

/* ------------------------------------------------------------------ model fits */
typedef struct {
    /* inputs */
    const double *y;        /* n_vox x m, C order */
    double *dirs;           /* n_vox x 3, flipped in place like the reference */
    const int16_t *htable;
    int64_t n_vox; int m, ndirs;
    const float *rot[2];    /* rotated LUT blocks, each (n_rot[b], ndirs, m) float32 C order */
    int n_rot[2];
    const float *iso;       /* (n_iso, m) float32 */
    int n_iso;
    double lambda1, lambda2;
    int flags;              /* bit0 rmse, bit1 nrmse, bit2 model extra (modulated / corrected DWI) */
    /* NODDI */
    const double *norms;    /* dwi_count x n_wm */
    const float *icvf, *kappa;
    const int64_t *dwi_idx; int dwi_count; int exvivo;
    /* FreeWater */
    int mouse;
    /* SANDI / CZB */
    const double *A_shared; /* m x n col-major (SANDI) */
    const double *sandi_norms, *Rs, *d_in, *d_isos;
    int n_rs, n_in;
    /* outputs */
    double *est; int n_maps;
    double *rmse, *nrmse, *extra;
    int32_t *lut_out;       /* optional: LUT index per voxel */
    int32_t *support_out;   /* optional NODDI: stage-2 support size */
    int *err_voxel;
} fit_args;

typedef struct { fit_args *a; int model; int64_t i0, i1; int status; } chunk_t;

static void assemble(const fit_args *a, int k, double *A)
{
    int m = a->m, b, j, r, col = 0;
    for (b = 0; b < 2; ++b)
        for (j = 0; j < a->n_rot[b]; ++j, ++col) {
            const float *src = a->rot[b] + ((size_t)j * a->ndirs + k) * m;
            for (r = 0; r < m; ++r) A[(size_t)col * m + r] = (double)src[r];
        }
    (void)col;
}

static int fit_noddi_range(fit_args *a, int64_t i0, int64_t i1)
{
    int m = a->m, n_wm = a->n_rot[0], n = n_wm + 1 + (a->exvivo ? 1 : 0), dc = a->dwi_count;
    int single_b0 = (m == 1 + dc);
    double *A = (double *)calloc((size_t)m * n, sizeof(double));
    double *A2 = (double *)calloc((size_t)dc * n_wm, sizeof(double));
    double *A3 = (double *)calloc((size_t)m * n, sizeof(double));
    double *y2 = (double *)calloc(dc, sizeof(double));
    double *x = (double *)calloc(n, sizeof(double)), *x3 = (double *)calloc(n, sizeof(double));
    double *yest = (double *)calloc(m, sizeof(double));
    int *pos = (int *)calloc(n, sizeof(int));
    nnls_ws nw; lars_ws lw;
    int64_t i; int j, k, r, status = 0;
    nnls_ws_init(&nw); lars_ws_init(&lw);
    for (i = i0; i < i1; ++i) {
        const double *y = a->y + (size_t)i * m;
        double rn, s_all, s_wm, f1, f2, k1, ndi, odi, fwf;
        int pc = 0;
        int lut = orc_dir_to_lut_idx(a->dirs + 3 * i, a->htable, NULL, NULL);
        if (lut < 0) { status = 1; if (a->err_voxel) *a->err_voxel = (int)i; break; }
        if (a->lut_out) a->lut_out[i] = lut;
        assemble(a, lut, A);
        if (a->exvivo) for (r = 0; r < m; ++r) A[(size_t)(n - 2) * m + r] = 1.0;
        for (r = 0; r < m; ++r) A[(size_t)(n - 1) * m + r] = (double)a->iso[r];
        /* fit 1: isotropic fraction */
        nnls_core(&nw, A, y, m, n, x, &rn);
        /* fit 2: support selection on the normalised DWI rows */
        for (j = 0; j < dc; ++j) {
            r = single_b0 ? j + 1 : (int)a->dwi_idx[j];
            for (k = 0; k < n_wm; ++k) A2[(size_t)k * dc + j] = A[(size_t)k * m + r] * a->norms[(size_t)j * n_wm + k];
            y2[j] = y[r] - x[n - 1] * (double)a->iso[r];
            if (a->exvivo) y2[j] = y2[j] - x[n - 2] * 1.0;
            if (y2[j] < 0.0) y2[j] = 0.0;
        }
        lars_core(&lw, A2, y2, dc, n_wm, x, a->lambda1, a->lambda2, 0);
        /* fit 3: debias on the support */
        if (a->exvivo) x[n - 2] = 1.0;
        x[n - 1] = 1.0;
        for (j = 0; j < n; ++j) if (x[j] > 0.0) pos[pc++] = j;
        if (a->support_out) a->support_out[i] = pc;
        for (k = 0; k < pc; ++k) memcpy(A3 + (size_t)k * m, A + (size_t)pos[k] * m, sizeof(double) * m);
        nnls_core(&nw, A3, y, m, pc, x3, &rn);
        for (j = 0; j < pc; ++j) x[pos[j]] = x3[j];
        /* maps */
        s_all = 0.0; s_wm = 0.0; f1 = f2 = k1 = 0.0;
        for (j = 0; j < n; ++j) s_all += x[j];
        s_all += 1e-16;
        for (j = 0; j < n_wm; ++j) s_wm += x[j] / s_all;
        s_wm += 1e-16;
        for (j = 0; j < n_wm; ++j) {
            f1 += a->icvf[j] * x[j] / s_all / s_wm;
            f2 += ((float)(1.0 - a->icvf[j])) * x[j] / s_all / s_wm;
            k1 += a->kappa[j] * x[j] / s_all / s_wm;
        }
        ndi = f1 / (f1 + f2 + 1e-16);
        odi = 2.0 / M_PI * atan2(1.0, k1);
        fwf = x[n - 1] / s_all;
        a->est[(size_t)i * a->n_maps + 0] = ndi;
        a->est[(size_t)i * a->n_maps + 1] = odi;
        a->est[(size_t)i * a->n_maps + 2] = fwf;
        if (a->exvivo) a->est[(size_t)i * a->n_maps + 3] = x[n - 2] / s_all;
        if (a->flags & 1) a->rmse[i] = fit_rmse(A, m, n, y, x, yest);
        if (a->flags & 2) a->nrmse[i] = fit_nrmse(A, m, n, y, x, yest);
        if (a->flags & 4) {
            double tf = 1.0 - fwf;
            a->extra[2 * i + 0] = ndi * tf;
            a->extra[2 * i + 1] = odi * tf;
        }
    }
    nnls_ws_free(&nw); lars_ws_free(&lw);
    free(A); free(A2); free(A3); free(y2); free(x); free(x3); free(yest); free(pos);
    return status;
}

static int fit_freewater_range(fit_args *a, int64_t i0, int64_t i1)
{
    int m = a->m, n_perp = a->n_rot[0], n_iso = a->n_iso, n = n_perp + n_iso;
    double *A = (double *)calloc((size_t)m * n, sizeof(double));
    double *x = (double *)calloc(n, sizeof(double)), *yest = (double *)calloc(m, sizeof(double));
    lars_ws lw; int64_t i; int j, k, r, status = 0;
    lars_ws_init(&lw);
    for (i = i0; i < i1; ++i) {
        const double *y = a->y + (size_t)i * m;
        double xs = 0.0, xp = 0.0, v;
        int lut = orc_dir_to_lut_idx(a->dirs + 3 * i, a->htable, NULL, NULL);
        if (lut < 0) { status = 1; if (a->err_voxel) *a->err_voxel = (int)i; break; }
        if (a->lut_out) a->lut_out[i] = lut;
        assemble(a, lut, A);
        for (j = 0; j < n_iso; ++j)
            for (r = 0; r < m; ++r) A[(size_t)(n_perp + j) * m + r] = (double)a->iso[(size_t)j * m + r];
        lars_core(&lw, A, y, m, n, x, a->lambda1, a->lambda2, 0);
        for (j = 0; j < n; ++j) { xs += x[j]; if (j < n_perp) xp += x[j]; }
        xs += 1e-16;
        v = xp / xs;
        a->est[(size_t)i * a->n_maps + 0] = v;
        a->est[(size_t)i * a->n_maps + 1] = 1.0 - v;
        if (a->mouse) {
            a->est[(size_t)i * a->n_maps + 2] = x[n_perp] / xs;
            a->est[(size_t)i * a->n_maps + 3] = x[n_perp + 1] / xs;
        }
        if (a->flags & 1) a->rmse[i] = fit_rmse(A, m, n, y, x, yest);
        if (a->flags & 2) a->nrmse[i] = fit_nrmse(A, m, n, y, x, yest);
        if (a->flags & 4) {
            for (j = 0; j < n - n_iso; ++j) x[j] = 0.0;
            for (j = 0; j < m; ++j) {
                double fw = 0.0, c;
                for (k = 0; k < n; ++k) fw += A[(size_t)k * m + j] * x[k];
                c = y[j] - fw;
                a->extra[(size_t)i * m + j] = c < 0.0 ? 0.0 : c;
            }
        }
    }
    lars_ws_free(&lw);
    free(A); free(x); free(yest);
    return status;
}

static int fit_czb_range(fit_args *a, int64_t i0, int64_t i1)
{
    int m = a->m, n_rs = a->n_rot[0], n_perp = a->n_rot[1], n_iso = a->n_iso, n = n_rs + n_perp + n_iso;
    double *A = (double *)calloc((size_t)m * n, sizeof(double));
    double *x = (double *)calloc(n, sizeof(double)), *yest = (double *)calloc(m, sizeof(double));
    lars_ws lw; int64_t i; int j, r, status = 0;
    lars_ws_init(&lw);
    for (i = i0; i < i1; ++i) {
        const double *y = a->y + (size_t)i * m;
        double f1 = 0.0, f2 = 0.0, aa = 0.0, v, d;
        int lut = orc_dir_to_lut_idx(a->dirs + 3 * i, a->htable, NULL, NULL);
        if (lut < 0) { status = 1; if (a->err_voxel) *a->err_voxel = (int)i; break; }
        if (a->lut_out) a->lut_out[i] = lut;
        assemble(a, lut, A);
        for (j = 0; j < n_iso; ++j)
            for (r = 0; r < m; ++r) A[(size_t)(n_rs + n_perp + j) * m + r] = (double)a->iso[(size_t)j * m + r];
        lars_core(&lw, A, y, m, n, x, a->lambda1, a->lambda2, 0);
        for (j = 0; j < n_rs + n_perp; ++j) {
            if (j < n_rs) f1 += x[j];
            if (j >= n_rs && j < n_rs + n_perp) f2 += x[j];
        }
        f2 += 1e-16;
        v = f1 / (f1 + f2 + 1e-16);
        f1 += 1e-16;
        for (j = 0; j < n_rs; ++j) aa += a->Rs[j] * x[j];
        aa = 1e6 * 2.0 * aa / f1;
        d = (4.0 * v) / (M_PI * pow(aa, 2.0) + 1e-16);
        a->est[(size_t)i * 3 + 0] = v;
        a->est[(size_t)i * 3 + 1] = aa;
        a->est[(size_t)i * 3 + 2] = d;
        if (a->flags & 1) a->rmse[i] = fit_rmse(A, m, n, y, x, yest);
        if (a->flags & 2) a->nrmse[i] = fit_nrmse(A, m, n, y, x, yest);
    }
    lars_ws_free(&lw);
    free(A); free(x); free(yest);
    return status;
}

static int fit_sandi_range(fit_args *a, int64_t i0, int64_t i1)
{
    int m = a->m, n_rs = a->n_rs, n_in = a->n_in, n_iso = a->n_iso, n = n_rs + n_in + n_iso;
    const double *A = a->A_shared;
    double *x = (double *)calloc(n, sizeof(double)), *yest = (double *)calloc(m, sizeof(double));
    lars_ws lw; int64_t i; int j;
    lars_ws_init(&lw);
    for (i = i0; i < i1; ++i) {
        const double *y = a->y + (size_t)i * m;
        double xs = 0, sph = 0, stk = 0, iso = 0, Rsoma = 0, Din = 0, De = 0;
        lars_core(&lw, A, y, m, n, x, a->lambda1, a->lambda2, 0);
        for (j = 0; j < n; ++j) x[j] = x[j] * a->sandi_norms[j];
        for (j = 0; j < n; ++j) {
            xs += x[j];
            if (j < n_rs) sph += x[j];
            if (j >= n_rs && j < n_rs + n_in) stk += x[j];
            if (j >= n_rs + n_in) iso += x[j];
        }
        xs += 1e-16;
        a->est[(size_t)i * 6 + 0] = sph / xs;
        a->est[(size_t)i * 6 + 1] = stk / xs;
        a->est[(size_t)i * 6 + 2] = iso / xs;
        for (j = 0; j < n; ++j) {
            if (j < n_rs) Rsoma += a->Rs[j] * x[j];
            if (j >= n_rs && j < n_rs + n_in) Din += a->d_in[j - n_rs] * x[j];
            if (j >= n_rs + n_in) De += a->d_isos[j - (n_rs + n_in)] * x[j];
        }
        sph += 1e-16; stk += 1e-16; iso += 1e-16;
        a->est[(size_t)i * 6 + 3] = 1e6 * Rsoma / sph;
        a->est[(size_t)i * 6 + 4] = 1e3 * Din / stk;
        a->est[(size_t)i * 6 + 5] = 1e3 * De / iso;
        /* reference quirk: errors use the normalised A with the re-scaled x (models.pyx:1570-1571, 1615) */
        if (a->flags & 1) a->rmse[i] = fit_rmse(A, m, n, y, x, yest);
        if (a->flags & 2) a->nrmse[i] = fit_nrmse(A, m, n, y, x, yest);
    }
    lars_ws_free(&lw);
    free(x); free(yest);
    return 0;
}

enum { ORC_NODDI = 0, ORC_FREEWATER = 1, ORC_CZB = 2, ORC_SANDI = 3 };

static void *chunk_main(void *p)
{
    chunk_t *c = (chunk_t *)p;
    switch (c->model) {
    case ORC_NODDI: c->status = fit_noddi_range(c->a, c->i0, c->i1); break;
    case ORC_FREEWATER: c->status = fit_freewater_range(c->a, c->i0, c->i1); break;
    case ORC_CZB: c->status = fit_czb_range(c->a, c->i0, c->i1); break;
    default: c->status = fit_sandi_range(c->a, c->i0, c->i1); break;
    }
    return NULL;
}

/* Contiguous voxel chunks, one per thread (BaseModel.fit, amico/models.pyx:204-211; the last chunk
 * absorbs the remainder).  Unlike the reference this also accepts n_vox < nthreads. */
static int run_chunks(fit_args *a, int model, int nthreads)
{
    int64_t n = a->n_vox, c;
    int t, nt, status = 0;
    chunk_t *ch; pthread_t *th;
    if (nthreads < 1) nthreads = 1;
    if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
    c = n / nthreads;
    nt = nthreads;
    ch = (chunk_t *)calloc(nt, sizeof(chunk_t));
    th = (pthread_t *)calloc(nt, sizeof(pthread_t));
    for (t = 0; t < nt; ++t) {
        ch[t].a = a; ch[t].model = model;
        ch[t].i0 = t * c; ch[t].i1 = (t == nt - 1) ? n : (t + 1) * c;
    }
    if (nt == 1) chunk_main(&ch[0]);
    else {
        for (t = 0; t < nt; ++t) pthread_create(&th[t], NULL, chunk_main, &ch[t]);
        for (t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    }
    for (t = 0; t < nt; ++t) if (ch[t].status) status = ch[t].status;
    free(ch); free(th);
    return status;
}

int orc_fit_noddi(const double *y, double *dirs, int64_t n_vox, int m, const int16_t *htable, int ndirs,
                  const float *wm, int n_wm, const float *iso, const double *norms, const float *icvf,
                  const float *kappa, const int64_t *dwi_idx, int dwi_count, int exvivo,
                  double lambda1, double lambda2, int flags, int nthreads,
                  double *est, double *rmse, double *nrmse, double *est_mod,
                  int32_t *lut_out, int32_t *support_out, int *err_voxel)
{
    fit_args a; memset(&a, 0, sizeof(a));
    a.y = y; a.dirs = dirs; a.n_vox = n_vox; a.m = m; a.htable = htable; a.ndirs = ndirs;
    a.rot[0] = wm; a.n_rot[0] = n_wm; a.iso = iso; a.n_iso = 1; a.norms = norms; a.icvf = icvf; a.kappa = kappa;
    a.dwi_idx = dwi_idx; a.dwi_count = dwi_count; a.exvivo = exvivo; a.lambda1 = lambda1; a.lambda2 = lambda2;
    a.flags = flags; a.est = est; a.n_maps = exvivo ? 4 : 3; a.rmse = rmse; a.nrmse = nrmse; a.extra = est_mod;
    a.lut_out = lut_out; a.support_out = support_out; a.err_voxel = err_voxel;
    return run_chunks(&a, ORC_NODDI, nthreads);
}

int orc_fit_freewater(const double *y, double *dirs, int64_t n_vox, int m, const int16_t *htable, int ndirs,
                      const float *D, int n_perp, const float *CSF, int n_iso, int mouse,
                      double lambda1, double lambda2, int flags, int nthreads,
                      double *est, double *rmse, double *nrmse, double *y_corrected,
                      int32_t *lut_out, int *err_voxel)
{
    fit_args a; memset(&a, 0, sizeof(a));
    a.y = y; a.dirs = dirs; a.n_vox = n_vox; a.m = m; a.htable = htable; a.ndirs = ndirs;
    a.rot[0] = D; a.n_rot[0] = n_perp; a.iso = CSF; a.n_iso = n_iso; a.mouse = mouse;
    a.lambda1 = lambda1; a.lambda2 = lambda2; a.flags = flags;
    a.est = est; a.n_maps = mouse ? 4 : 2; a.rmse = rmse; a.nrmse = nrmse; a.extra = y_corrected;
    a.lut_out = lut_out; a.err_voxel = err_voxel;
    return run_chunks(&a, ORC_FREEWATER, nthreads);
}

int orc_fit_czb(const double *y, double *dirs, int64_t n_vox, int m, const int16_t *htable, int ndirs,
                const float *wmr, int n_rs, const float *wmh, int n_perp, const float *iso, int n_iso,
                const double *Rs, double lambda1, double lambda2, int flags, int nthreads,
                double *est, double *rmse, double *nrmse, int32_t *lut_out, int *err_voxel)
{
    fit_args a; memset(&a, 0, sizeof(a));
    a.y = y; a.dirs = dirs; a.n_vox = n_vox; a.m = m; a.htable = htable; a.ndirs = ndirs;
    a.rot[0] = wmr; a.n_rot[0] = n_rs; a.rot[1] = wmh; a.n_rot[1] = n_perp; a.iso = iso; a.n_iso = n_iso;
    a.Rs = Rs; a.lambda1 = lambda1; a.lambda2 = lambda2; a.flags = flags;
    a.est = est; a.n_maps = 3; a.rmse = rmse; a.nrmse = nrmse; a.lut_out = lut_out; a.err_voxel = err_voxel;
    return run_chunks(&a, ORC_CZB, nthreads);
}

int orc_fit_sandi(const double *y, int64_t n_vox, int m, const double *signal, const double *norms,
                  const double *Rs, int n_rs, const double *d_in, int n_in, const double *d_isos, int n_iso,
                  double lambda1, double lambda2, int flags, int nthreads,
                  double *est, double *rmse, double *nrmse)
{
    fit_args a; memset(&a, 0, sizeof(a));
    a.y = y; a.n_vox = n_vox; a.m = m; a.A_shared = signal; a.sandi_norms = norms;
    a.Rs = Rs; a.n_rs = n_rs; a.d_in = d_in; a.n_in = n_in; a.d_isos = d_isos; a.n_iso = n_iso;
    a.lambda1 = lambda1; a.lambda2 = lambda2; a.flags = flags;
    a.est = est; a.n_maps = 6; a.rmse = rmse; a.nrmse = nrmse;
    return run_chunks(&a, ORC_SANDI, nthreads);
}

/* Batch helper for the LUT-index parity tests: idx[i] = lut(dirs[i]) (dirs flipped in place). */
int orc_lut_indices(double *dirs, int64_t n, const int16_t *htable, int32_t *idx)
{
    int64_t i; int bad = 0;
    for (i = 0; i < n; ++i) {
        idx[i] = orc_dir_to_lut_idx(dirs + 3 * i, htable, NULL, NULL);
        if (idx[i] < 0) bad = 1;
    }
    return bad;
}
