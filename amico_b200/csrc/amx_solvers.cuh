// Warp-per-voxel solvers.  One warp owns one voxel; lane l owns atoms l, l+32, ... (NPL per lane)
// and active-set position l.  Per-direction Gram tables live in global memory (L2-resident).
//
//  warp_lars : the `lasso` the reference calls (cyspams.interfaces.lasso -- SPAMS LARS/homotopy,
//              mode PENALTY, pos=true; call sites amico/models.pyx:615, 926, 1238, 1569).  Same
//              path, same operation order and un-fused arithmetic as the CPU restatement in
//              oracle/amico_oracle.c::lars_core, so results agree bit for bit given equal inputs.
//  warp_nnls : the `nnls` the reference calls (Lawson-Hanson; amico/models.pyx:911, 940) run in
//              Gram space: same pivoting (largest dual enters, ratio test leaves, candidate
//              independence + positivity test) with the passive-set system solved through a
//              Cholesky factor of H_PP instead of a Householder QR of A_P.
#pragma once
#include "amx_warp.cuh"
#ifndef AMX_GD
#define AMX_GD 2  // Gram rows fetched per batch (measured at 64-80 registers: 1 -> 39.6 ms, 2 -> 37.7, 3 -> 39.1, 4 -> 38.9) in the NODDI solvers' O(n |P|) passes (NPL > 2)
#endif
#include <math.h>
#ifndef AMX_SOLVER_INLINE
#define AMX_SOLVER_INLINE __noinline__
#endif

namespace amx {

// ------------------------------------------------------------------------------------------------
// Triangular solves against the packed lower factor Lp (row-major packed).  Lane a holds element a of the right-hand
// side / the result (0 beyond np) and rdl = 1 / L[a][a] of its own row, so the pivot of step k is formed on lane k before
// the broadcast and no per-step select or reciprocal load is needed.
__device__ __forceinline__ double fwd_subst(const double *Lp, double rdl, int np, double t, int lane)
{
    const double *row = Lp + tri(lane, 0);
    const bool act = lane < np;
    #pragma unroll 1
    for (int k = 0; k < np; ++k) {
        const double vk = shfl(t * rdl, k);  // lane k's t is final from step k - 1 on
        if (act && lane > k) t = fma(-row[k], vk, t);
    }
    return act ? t * rdl : 0.0;
}

__device__ __forceinline__ double back_subst(const double *Lp, double rdl, int np, double t, int lane)
{
    const double *col = Lp + tri(np - 1, 0) + lane;  // element (k, lane) of row k, walking up
    #pragma unroll 1
    for (int k = np - 1; k >= 0; --k) {
        const double sk = shfl(t * rdl, k);
        if (lane < k) t = fma(-*col, sk, t);
        col -= k;
    }
    return lane < np ? t * rdl : 0.0;
}

struct NnlsStat {
    int outer, inner, removed;
};

// What warp_nnls needs to look at a candidate column in A-space (optional): the direction's dictionary slab S (row-major
// [m][n_pad], fp32-valued), and the voxel's signal.
struct ASpace {
    const float *S; int n_pad, m;
    const void *y; int y_f64; long long vox;
};

// arg-max over strictly positive values (lanes with idx < 0 do not take part): the bit pattern of a positive double orders like
// an unsigned integer, so no order-preserving key is needed; ties -> lowest index
__device__ __forceinline__ void warp_argmax_pos(double &v, int &idx)
{
    const unsigned hi = idx >= 0 ? (unsigned)__double2hiint(v) : 0u, lo = idx >= 0 ? (unsigned)__double2loint(v) : 0u;
    const unsigned hm = __reduce_max_sync(FULL, hi);
    const unsigned lm = __reduce_max_sync(FULL, hi == hm ? lo : 0u);
    const bool win = idx >= 0 && hi == hm && lo == lm;
    const int widx = __reduce_min_sync(FULL, win ? idx : 0x7fffffff);
    v = __hiloint2double((int)hm, (int)lm);
    idx = (hm | lm) ? widx : -1;
}

// Cholesky downdate for the deletion of passive position q (of pn): rows q+1.. move up one slot and Givens rotations restore
// the triangle; z = L^-1 c_P is rotated along (the reference updates its QR the same way).  Row i = lane streams its OLD row
// i + 1 through the rotations: per step one load, one store, the rotated-out element carried in a register -- no separate
// shift pass for the columns >= q.  rdl / zl: per-lane 1 / diagonal and z, updated in place.
__device__ __forceinline__ void chol_delete(double *Lp, int q, int pn, double &rdl, double &zl, int lane)
{
    const bool mine = (lane >= q) && (lane < pn - 1);
    const double *src = Lp + tri(lane + 1, 0);  // old row lane + 1 (only dereferenced when `mine`)
    double *dst = Lp + tri(lane, 0);
    #pragma unroll 1
    for (int col = 0; col < q; ++col) {  // columns left of q: plain move (read everywhere before anyone overwrites)
        double lv = 0.0;
        if (mine) lv = src[col];
        __syncwarp();
        if (mine) dst[col] = lv;
    }
    double carry = 0.0;
    if (mine) carry = src[q];
    __syncwarp();
    #pragma unroll 1
    for (int r = q; r < pn - 1; ++r) {
        const bool act = mine && lane >= r;
        double u2 = 0.0;
        if (act) u2 = src[r + 1];
        const double a = shfl(carry, r), b = shfl(u2, r);
        const double ir = rsqrt(fma(a, a, b * b));  // = 1 / (new diagonal element)
        const double cs = a * ir, sn = b * ir;
        __syncwarp();  // column r of row lane + 1 was read one step ago; its owner may overwrite it now
        if (act) {
            dst[r] = fma(cs, carry, sn * u2);
            carry = fma(cs, u2, -sn * carry);
            if (lane == r) rdl = ir;
        }
        const double zr = shfl(zl, r), zr1 = shfl(zl, r + 1);
        if (lane == r) zl = fma(cs, zr, sn * zr1);
        else if (lane == r + 1) zl = fma(cs, zr1, -sn * zr);
    }
    if (lane >= pn - 1) zl = 0.0;
    __syncwarp();
}

// min 1/2 x'Tx - c'x, x >= 0 over the atoms whose bit is set in `allowed` (bit s of lane l <-> atom
// l + 32 s).  T: n x n Gram (ld ldT), c/x: per-warp shared arrays (c readable up to 32 NPL entries).  mcap = number of rows of the
// least-squares system (the reference stops growing the passive set at m).  Returns overflow flag.
// MAPPED (NPL must be 1): the system is the sub-system of T on the atoms map[0..n) (n <= 32, one per lane): c, x, P
// live in that compact numbering and T is addressed through map -- NODDI stage 3, where the support holds ~10 of 145
// atoms, so the dual pass touches one Gram entry per lane and row instead of NPL.
// zz_out (optional): ||z||^2 = ||A x||^2 of the final passive system.
template <int NPL, bool MAPPED = false>
__device__ AMX_SOLVER_INLINE int warp_nnls(const double *__restrict__ T, int ldT, int n, int mcap, int itmax, const double *c, double *x,
                         unsigned allowed, double *Lp, double *, int *P, int lane, NnlsStat *st, int cap = LC,
                         const int *map = nullptr, const ASpace *as = nullptr, double *zz_out = nullptr)
{
    static_assert(!MAPPED || NPL == 1, "the mapped variant keeps one atom per lane");
    auto AT = [&](int q) { return MAPPED ? map[q] : q; };  // compact index -> atom (= Gram table row / column)
    const int mycol = MAPPED ? map[lane < n ? lane : 0] : 0;
    int np = 0, iter = 0, overflow = 0;
    cap = min(cap, c_lc_cap);
    unsigned inP = 0, avail = 0;
    double xp = 0.0, zl = 0.0, rdl = 0.0;  // coefficient, z and 1 / diagonal of this lane's passive position
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        x[lane + 32 * s] = 0.0;
        avail |= (lane + 32 * s < n ? 1u : 0u) << s;
    }
    avail &= allowed;
    __syncwarp();
    for (;;) {
        if (np >= mcap) break;
        if (np >= cap) { overflow = 1; break; }
        // dual w = c - T[:,P] x_P; slots outside `valid` carry finite values nobody looks at
        double wl[NPL];
        unsigned valid = avail & ~inP;
#pragma unroll
        for (int s = 0; s < NPL; ++s) wl[s] = c[lane + 32 * s];
        {   // rows of T are L2-resident (~300 cycles): fetch GD rows at a time
            constexpr int GD = (NPL <= 2) ? 4 : AMX_GD;
#pragma unroll 1
            for (int k0 = 0; k0 < np; k0 += GD) {
                double gq[GD][NPL];
#pragma unroll
                for (int q = 0; q < GD; ++q) {
                    const double *row = T + (size_t)AT(P[min(k0 + q, np - 1)]) * ldT;
#pragma unroll
                    for (int s = 0; s < NPL; ++s) gq[q][s] = row[MAPPED ? mycol : lane + 32 * s];  // table is padded
                }
#pragma unroll
                for (int q = 0; q < GD; ++q) {
                    if (k0 + q < np) {
                        const double xk = x[P[k0 + q]];
#pragma unroll
                        for (int s = 0; s < NPL; ++s) wl[s] = fma(-gq[q][s], xk, wl[s]);
                    }
                }
            }
        }
        // candidate selection
        int j = -1;
        double v = 0.0, d2 = 0.0, znum = 0.0;
        for (;;) {
            double bv = 0.0;
            int bj = -1;
#pragma unroll
            for (int s = 0; s < NPL; ++s)
                if (((valid >> s) & 1u) && wl[s] > bv) { bv = wl[s]; bj = lane + 32 * s; }
            warp_argmax_pos(bv, bj);
            if (bj < 0) { j = -1; break; }
            j = bj;
            double t = (lane < np) ? T[(size_t)AT(j) * ldT + AT(P[lane])] : 0.0;  // = T[P[lane]][j] (the table is exactly symmetric): one row, not np
            v = fwd_subst(Lp, rdl, np, t, lane);
            double vv = v * v, vz = v * zl;  // both 0 beyond np (v is); two interleaved butterfly sums
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vv += shfl_xor(vv, o);
                vz += shfl_xor(vz, o);
            }
            const double hjj = T[(size_t)AT(j) * (ldT + 1)];
            d2 = hjj - vv;
            znum = c[j] - vz;
            if (as && np > 0 && d2 < 1e-10 * hjj) {
                // Near-dependent candidate: H_jj - v.v has lost its digits (the Gram form squares the conditioning; below ~1e-13 H_jj
                // it is rounding noise of either sign) and so has c_j - v.z -- yet the NODDI dictionary really holds atoms that are
                // independent of the passive set only at the 1e-7 level (d2 ~ 1e-14 H_jj), and the reference's Householder QR
                // (amico/models.pyx:911, 940 -> Lawson-Hanson) resolves and accepts them.  Re-evaluate both quantities in A-space:
                // r = a_j - A_P beta with beta = L^-T v is the part of the column orthogonal to the passive set, d2 = |r|^2 and the
                // numerator is r.y.  Measured on the C model of this solver (tools/research): support mismatches against the
                // oracle 68 -> 0 of 6000 voxels at SNR 300, 8 -> 0 at SNR 30, for any threshold between 1e-6 and 1e-13; 1e-10 sends
                // ~0.3 candidates per voxel here.
                const double beta = back_subst(Lp, rdl, np, v, lane);
                const float *Sj = as->S + AT(j);
                double a2 = 0.0, ay = 0.0;
                #pragma unroll 1
                for (int i0 = 0; i0 < as->m; i0 += 32) {
                    const int i = i0 + lane;
                    const bool on = i < as->m;
                    const float *Si = as->S + (size_t)(on ? i : 0) * as->n_pad;
                    double r = (double)Sj[(size_t)(on ? i : 0) * as->n_pad];
                    #pragma unroll 1
                    for (int a = 0; a < np; a += 2) {  // two passive atoms per trip: both dictionary loads in flight together
                        const int a1 = min(a + 1, np - 1);
                        const float s0 = Si[AT(P[a])], s1 = Si[AT(P[a1])];
                        const double b0 = shfl(beta, a), b1 = (a + 1 < np) ? shfl(beta, a1) : 0.0;
                        r = fma(-(double)s0, b0, r);
                        r = fma(-(double)s1, b1, r);
                    }
                    if (on) {
                        const double yi = as->y_f64 ? ((const double *)as->y)[as->vox * as->m + i] : (double)((const float *)as->y)[as->vox * as->m + i];
                        a2 = fma(r, r, a2);
                        ay = fma(r, yi, ay);
                    }
                }
                d2 = warp_sum(a2);
                znum = warp_sum(ay);
                // r carries an absolute error of a few ulps of |a_j| (1 + |beta|_1): below 1e-12 |a_j| it is noise, the column counts
                // as dependent (the reference's own test rejects at 1.1e-14 |a_j| on a QR that is accurate to ~1e-16 |a_j|)
                if (d2 < 1e-24 * hjj) d2 = 0.0;
            }
            if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) break;
            // reject: drop j from this round's candidates
            if ((j & 31) == lane) valid &= ~(1u << (j >> 5));
        }
        if (j < 0) break;
        // move j to the passive set: append a row to the factor
        {
            const double ird = rsqrt(d2), dd = d2 * ird;
            const double znew = znum * ird;
            if (lane < np) Lp[tri(np, lane)] = v;
            if (lane == np) {
                Lp[tri(np, np)] = dd;
                rdl = ird;
                P[np] = j;
                zl = znew;
                xp = 0.0;
            }
            if ((j & 31) == lane) inP |= 1u << (j >> 5);
            ++np;
            if (st) ++st->outer;
        }
        __syncwarp();
        // secondary loop
        double s = 0.0;
        for (;;) {
            if (++iter > itmax) goto done;
            if (st) ++st->inner;
            s = back_subst(Lp, rdl, np, (lane < np) ? zl : 0.0, lane);
            bool neg = (lane < np) && (s <= 0.0);
            if (!__any_sync(FULL, neg)) break;
            double tmin = INFINITY;
            int cand = -1;
            if (neg) {
                double tt = -xp / (s - xp);
                if (tt < 2.0) { tmin = tt; cand = lane; }
            }
            warp_argmin<true>(tmin, cand);
            if (cand < 0) break;
            if (lane < np) xp = fma(tmin, s - xp, xp);
            if (lane == cand) xp = 0.0;
            const bool keep = (lane < np) && (xp > 0.0);
            const unsigned kmask = __ballot_sync(FULL, keep);
            const int np_old = np;
            unsigned rmask = ~kmask & (np_old >= 32 ? 0xffffffffu : ((1u << np_old) - 1u));  // removed positions
            const int myP = (lane < np) ? P[lane] : 0;
            if (lane < np && !keep) x[myP] = 0.0;
            for (unsigned r2 = rmask; r2; r2 &= r2 - 1) {  // their atoms leave the passive bit set
                const int a = __shfl_sync(FULL, myP, __ffs(r2) - 1);
                if ((a & 31) == lane) inP &= ~(1u << (a >> 5));
            }
            const int nnew = __popc(kmask);
            if (st) st->removed += np - nnew;
            const unsigned src = __fns(kmask, 0, lane + 1);
            const bool has = lane < nnew;
            const int srcl = has ? (int)src : 0;
            const double xs = shfl(xp, srcl);
            const int ps = __shfl_sync(FULL, myP, srcl);
            __syncwarp();
            if (has) P[lane] = ps;
            xp = has ? xs : 0.0;
            np = nnew;
            __syncwarp();
            if (np == 0) break;
            // Cholesky downdate (column deletion), one removed position at a time, highest first
            for (int pn = np_old; rmask; --pn) {
                const int q = 31 - __clz(rmask);
                rmask &= ~(1u << q);
                chol_delete(Lp, q, pn, rdl, zl, lane);
            }
        }
        if (lane < np) {
            xp = s;
            x[P[lane]] = s;
        }
        __syncwarp();
    }
done:
    if (lane < np) x[P[lane]] = xp;
    // ||A x||^2 = ||z||^2 of the final passive system (z = L^-1 c_P): with ||y||^2 it gives the fit residual without touching A
    if (zz_out) *zz_out = warp_sum(lane < np ? zl * zl : 0.0);
    __syncwarp();
    return overflow;
}

// ------------------------------------------------------------------------------------------------
// Non-negative LARS on T = G + ridge*I (ridge = max(lambda2, 1e-10), baked into the diagonal), correlations DtR (destroyed),
// following oracle/amico_oracle.c::lars_core step by step.  Ltrue = min(rows, K) of the underlying
// least-squares system.  x: per-warp shared output (exact zeros off-support).
// Mi: packed upper inverse of G_SS; u, gs: LC doubles; ind: LC ints.
__device__ __forceinline__ double sym_at(const double *Mi, int r, int c) { return Mi[r <= c ? tri(c, r) : tri(r, c)]; }

template <int NPL>
__device__ __noinline__ int warp_lars(const double *__restrict__ T, int ldT, double ridge_in, int K, int Ltrue, double lambda1,
                         double *DtR, double normX, double *Mi, double *u, double *gs, int *ind, double *x, int lane,
                         int *steps_out, int cap = LC)
{
    cap = min(cap, c_lc_cap);
    (void)ridge_in;  // T already holds G + max(lambda2, 1e-10) I (k_set_ridge)
    int L = Ltrue < K ? Ltrue : K;
    int overflow = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) x[lane + 32 * s] = 0.0;
    __syncwarp();
    if (steps_out) *steps_out = 0;
    if (L <= 0) return 0;
    int cur;
    {
        double bv = 0.0;
        int bi = -1;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int k = lane + 32 * s;
            if (k < K) {
                double v = DtR[k];
                if (bi < 0 || v > bv) { bv = v; bi = k; }
            }
        }
        warp_argmax(bv, bi);
        if (fabs(bv) < lambda1) return 0;
        cur = bi;
    }
    int newAtom = 1, iter = 0, na = 0;
    double coef_l = 0.0;
    int ind_l = -1;
    unsigned act = 0;
    const int length_path = 4 * L;
    #pragma unroll 1
    for (int i = 0; i < L; ++i) {
        if (i < 0) break;  // the CPU path would read ind[-1] here; cannot happen with pos=true
        ++iter;
        if (newAtom) {
            if (i >= cap) { overflow = 1; na = i; break; }
            if (lane == i) { ind_l = cur; coef_l = 0.0; ind[i] = cur; }
            if ((cur & 31) == lane) act |= 1u << (cur >> 5);
            __syncwarp();
            double g = 0.0;
            if (lane <= i) {
                g = T[(size_t)cur * ldT + ind_l];
                gs[lane] = g;
            }
            __syncwarp();
            if (i == 0) {
                if (lane == 0) Mi[0] = 1.0 / g;
            } else {
                double ur = 0.0;
                if (lane < i) {
                    #pragma unroll 1
                    for (int c = 0; c < i; ++c) ur = madd(ur, sym_at(Mi, lane, c), gs[c]);
                    u[lane] = ur;
                }
                __syncwarp();
                double dot = 0.0;
                #pragma unroll 1
                for (int j = 0; j < i; ++j) dot = madd(dot, u[j], gs[j]);
                double schur = 1.0 / __dsub_rn(gs[i], dot);
                if (lane < i) {
                    double su = __dmul_rn(schur, ur);
                    #pragma unroll 1
                    for (int k = lane; k < i; ++k) Mi[tri(k, lane)] = __dadd_rn(Mi[tri(k, lane)], __dmul_rn(su, u[k]));
                    Mi[tri(i, lane)] = __dmul_rn(-schur, ur);
                }
                if (lane == i) Mi[tri(i, i)] = schur;
            }
            __syncwarp();
        }
        na = i + 1;
        // path direction u = invGs * sign(DtR_S)
        if (lane <= i) gs[lane] = DtR[ind_l] > 0.0 ? 1.0 : -1.0;
        __syncwarp();
        double ul = 0.0;
        if (lane <= i) {
            #pragma unroll 1
            for (int c = 0; c <= i; ++c) ul = madd(ul, sym_at(Mi, lane, c), gs[c]);
            u[lane] = ul;
        }
        __syncwarp();
        // largest step before an active coefficient crosses zero (last index wins ties)
        double step_max = INFINITY;
        int fz = -1;
        if (lane <= i) {
            double r = -coef_l / ul;
            if (r > 0.0) { step_max = r; fz = lane; }
        }
        warp_argmin<false>(step_max, fz);
        if (fz < 0) step_max = INFINITY;
        const double cc = fabs(DtR[ind[0]]);
        // correlation slopes  (T + ridge I)[:, S] u
        double sl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) sl[s] = 0.0;
        // The Gram rows are L2-resident (~300 cycles): fetch GD rows at a time, then accumulate them in path order.
        constexpr int GD = (NPL <= 2) ? 4 : 3;
#pragma unroll 1
        for (int j0 = 0; j0 <= i; j0 += GD) {
            double gq[GD][NPL];
            int aq[GD];
#pragma unroll
            for (int q = 0; q < GD; ++q) {
                aq[q] = ind[min(j0 + q, i)];
                const double *row = T + (size_t)aq[q] * ldT;
#pragma unroll
                for (int s = 0; s < NPL; ++s) gq[q][s] = (lane + 32 * s < K) ? row[lane + 32 * s] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < GD; ++q) {
                if (j0 + q <= i) {
                    const double uj = u[j0 + q];
#pragma unroll
                    for (int s = 0; s < NPL; ++s) {
                        if (lane + 32 * s < K) sl[s] = madd(sl[s], gq[q][s], uj);
                    }
                }
            }
        }
        // first inactive atom reaching the common correlation: entry of smallest magnitude, lowest index
        double tl[NPL];
        double bt = INFINITY;
        int bk = -1;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int k = lane + 32 * s;
            tl[s] = INFINITY;
            if (k < K) {
                if (!((act >> s) & 1u) && sl[s] < 1.0) tl[s] = __ddiv_rn(__dsub_rn(cc, DtR[k]), __dsub_rn(1.0, sl[s]));
                double at = fabs(tl[s]);
                if (bk < 0 || at < bt) { bt = at; bk = k; }
            }
        }
        warp_argmin<true>(bt, bk);
        double step;
        {
            double mine = 0.0;
#pragma unroll
            for (int s = 0; s < NPL; ++s)
                if (s == (bk >> 5)) mine = tl[s];
            step = shfl(mine, bk & 31);
        }
        cur = bk;
        double coeff1 = 0.0, coeff2 = 0.0;
        #pragma unroll 1
        for (int j = 0; j <= i; ++j) {
            double uj = u[j];
            coeff1 = __dadd_rn(coeff1, DtR[ind[j]] > 0.0 ? uj : -uj);
        }
        #pragma unroll 1
        for (int j = 0; j <= i; ++j) coeff2 = madd(coeff2, DtR[ind[j]], u[j]);
        const double step_max2 = __dsub_rn(cc, lambda1);
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        if (lane <= i) {
            coef_l = madd(coef_l, step, ul);
            if (coef_l < 0.0) coef_l = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int k = lane + 32 * s;
            if (k < K) DtR[k] = __dsub_rn(DtR[k], __dmul_rn(step, sl[s]));
        }
        normX = __dadd_rn(normX, __dsub_rn(__dmul_rn(__dmul_rn(coeff1, step), step), __dmul_rn(__dmul_rn(2.0, coeff2), step)));
        __syncwarp();
        if (step == step_max) {
            // remove active position z: shrink ind/coeffs and downdate the inverse
            const int z = fz;
            const int az = ind[z];
            const double schur_r = Mi[tri(z, z)];
            double uk = 0.0;
            if (lane < i) uk = (lane < z) ? Mi[tri(z, lane)] : Mi[tri(lane + 1, z)];
            __syncwarp();
            if (lane < i) u[lane] = uk;
            double cn = __shfl_down_sync(FULL, coef_l, 1);
            int in_ = __shfl_down_sync(FULL, ind_l, 1);
            if (lane >= z && lane < i) { coef_l = cn; ind_l = in_; }
            if (lane == i) { coef_l = 0.0; ind_l = -1; }
            if ((az & 31) == lane) act &= ~(1u << (az >> 5));
            #pragma unroll 1
            for (int j = z; j < i; ++j) {  // new column j <- old column j+1 without row z
                double mv = 0.0;
                if (lane <= j) mv = Mi[tri(j + 1, lane < z ? lane : lane + 1)];
                __syncwarp();
                if (lane <= j) Mi[tri(j, lane)] = mv;
                __syncwarp();
            }
            if (lane <= i) ind[lane] = ind_l;
            __syncwarp();
            if (lane < i)
                #pragma unroll 1
                for (int k = lane; k < i; ++k)
                    Mi[tri(k, lane)] = __dsub_rn(Mi[tri(k, lane)], __ddiv_rn(__dmul_rn(uk, u[k]), schur_r));
            __syncwarp();
            newAtom = 0;
            na = i;
            i -= 2;
        } else {
            newAtom = 1;
        }
        if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == L - 1) break;
    }
    if (lane < na && ind_l >= 0) x[ind_l] = coef_l;
    __syncwarp();
    if (steps_out) *steps_out = iter;
    return overflow;
}

// ------------------------------------------------------------------------------------------------
// warp_lars_fast: the same homotopy path as warp_lars (same entering / leaving rules, same stopping rules) written for
// throughput instead of bit-for-bit equality with the CPU arithmetic: fused multiply-adds, warp-shuffle reductions for
// the small sums, reciprocal instead of division for the candidate steps.  Used by NODDI stage 2, where only the SUPPORT
// of the (unique) elastic-net minimiser is consumed (amico/models.pyx:929-936) and its input already differs from the
// CPU's at the 1e-12 level through stage 1.
// sum_{c < n} M(r, c) v[c] for the packed symmetric matrix M (element (a, b), a >= b, at tri(a, b)); c runs uniformly over
// the warp, the index needs one select: (c < r) ? tri(r, 0) + c : tri(c, 0) + r
__device__ __forceinline__ double sym_row_dot(const double *Mi, int r, int n, const double *v)
{
    double acc = 0.0;
    const int base = tri(r, 0);
    int tc = 0;  // tri(c, 0)
    #pragma unroll 1
    for (int c = 0; c < n; ++c) {
        acc = fma(Mi[c < r ? base + c : tc + r], v[c], acc);
        tc += c + 1;
    }
    return acc;
}

template <int NPL>
__device__ __noinline__ int warp_lars_fast(const double *__restrict__ T, int ldT, int K, int Ltrue, double lambda1, double *DtR,
                                           double normX, double *Mi, double *u, double *gs, int *ind, double *x, int lane, int cap,
                                           unsigned (&sup)[NPL])
{
    // Result: sup[s] bit l set <=> atom l + 32 s ends with a positive coefficient (all lanes hold the words); when x is
    // not NULL the coefficients are also written there (exact zeros off-support).
    cap = min(cap, c_lc_cap);
    int L = Ltrue < K ? Ltrue : K;
    int overflow = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) sup[s] = 0u;
    if (x) {
#pragma unroll
        for (int s = 0; s < NPL; ++s) x[lane + 32 * s] = 0.0;
        __syncwarp();
    }
    if (L <= 0) return 0;
    int cur;
    {
        double bv = 0.0;
        int bi = -1;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int k = lane + 32 * s;
            if (k < K) {
                double v = DtR[k];
                if (bi < 0 || v > bv) { bv = v; bi = k; }
            }
        }
        warp_argmax(bv, bi);
        if (fabs(bv) < lambda1) return 0;
        cur = bi;
    }
    int newAtom = 1, iter = 0, na = 0;
    double coef_l = 0.0, rs_l = 0.0;  // coefficient / row sum of (G_SS)^-1 of this lane's active position
    int ind_l = -1;
    unsigned act = 0;
    const int length_path = 4 * L;
#pragma unroll 1
    for (int i = 0; i < L; ++i) {
        if (i < 0) break;
        ++iter;
        if (newAtom) {
            if (i >= cap) { overflow = 1; na = i; break; }
            if (lane == i) { ind_l = cur; coef_l = 0.0; ind[i] = cur; }
            if ((cur & 31) == lane) act |= 1u << (cur >> 5);
            __syncwarp();
            double g = 0.0;
            if (lane <= i) {
                g = T[(size_t)cur * ldT + ind_l];
                gs[lane] = g;
            }
            __syncwarp();
            if (i == 0) {
                if (lane == 0) Mi[0] = 1.0 / g;
                rs_l = (lane == 0) ? 1.0 / g : 0.0;
            } else {
                double ur = 0.0;
                if (lane < i) {
                    ur = sym_row_dot(Mi, lane, i, gs);
                    u[lane] = ur;
                }
                double dot = lane < i ? ur * g : 0.0, usum = lane < i ? ur : 0.0;  // two interleaved butterfly sums
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dot += shfl_xor(dot, o);
                    usum += shfl_xor(usum, o);
                }
                const double schur = 1.0 / (shfl(g, i) - dot);
                // row sums of the inverse after the Schur update: old rows += schur u_r (sum(u) - 1), new row = schur (1 - sum(u))
                if (lane < i) rs_l = fma(schur * ur, usum - 1.0, rs_l);
                if (lane == i) rs_l = schur * (1.0 - usum);
                __syncwarp();
                if (lane < i) {
                    const double su = schur * ur;
#pragma unroll 1
                    for (int k = lane; k < i; ++k) Mi[tri(k, lane)] = fma(su, u[k], Mi[tri(k, lane)]);
                    Mi[tri(i, lane)] = -su;
                }
                if (lane == i) Mi[tri(i, i)] = schur;
            }
            __syncwarp();
        }
        na = i + 1;
        // path direction u = invGs * sign(DtR_S)
        double dl = 0.0, sg = 0.0;
        if (lane <= i) {
            dl = DtR[ind_l];
            sg = dl > 0.0 ? 1.0 : -1.0;
            gs[lane] = sg;
        }
        __syncwarp();
        // With the positivity constraint every active correlation is positive (never observed otherwise: 0 of 150,000 model voxels),
        // so u = (G_SS)^-1 1 is the vector of row sums of the inverse, which is carried along in O(1) per lane and step; the
        // general form stays as the fallback.
        double ul = 0.0;
        if (!__any_sync(FULL, lane <= i && !(dl > 0.0))) {
            if (lane <= i) {
                ul = rs_l;
                u[lane] = ul;
            }
        } else if (lane <= i) {
            ul = sym_row_dot(Mi, lane, i + 1, gs);
            u[lane] = ul;
        }
        __syncwarp();
        // largest step before an active coefficient crosses zero (last index wins ties)
        double step_max = INFINITY;
        int fz = -1;
        // r = -coef / u is positive only for a positive coefficient that decreases (u < 0): most steps have none, and the fp64
        // division is a ~30-instruction subroutine the whole warp would walk through -- skip it (and the arg-min) then
        if (__any_sync(FULL, lane <= i && coef_l > 0.0 && ul < 0.0)) {
            if (lane <= i) {
                double r = -coef_l / ul;
                if (r > 0.0) { step_max = r; fz = lane; }
            }
            warp_argmin<false>(step_max, fz);
            if (fz < 0) step_max = INFINITY;
        }
        const double cc = fabs(shfl(dl, 0));
        // correlation slopes T[:, S] u; rows are L2-resident: fetch GD rows at a time
        double sl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) sl[s] = 0.0;
        constexpr int GD = (NPL <= 2) ? 4 : AMX_GD;
#pragma unroll 1
        for (int j0 = 0; j0 <= i; j0 += GD) {
            double gq[GD][NPL];
#pragma unroll
            for (int q = 0; q < GD; ++q) {
                const double *row = T + (size_t)ind[min(j0 + q, i)] * ldT + lane;
#pragma unroll
                for (int s = 0; s < NPL; ++s) gq[q][s] = row[32 * s];  // columns >= K: finite values of the next row / the padding,
                                                                       // only ever combined into slots that are masked by k < K
            }
#pragma unroll
            for (int q = 0; q < GD; ++q) {
                const double uj = (j0 + q <= i) ? u[j0 + q] : 0.0;
#pragma unroll
                for (int s = 0; s < NPL; ++s) sl[s] = fma(gq[q][s], uj, sl[s]);
            }
        }
        // first inactive atom reaching the common correlation: entry of smallest magnitude, lowest index.  Each lane first
        // picks the best of its own atoms by cross-multiplication (|a/b| < |c/d| <=> |a| d < |c| b for b, d > 0), so only one
        // reciprocal per lane and step is needed.
        double bnum = 0.0, bden = 0.0;  // best candidate of this lane: step = bnum / bden (bden > 0), none while bden == 0
        int mk = -1;                    // its atom; lanes without a candidate still offer their lowest atom (step = inf)
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            const int k = lane + 32 * s;
            if (k < K) {
                if (mk < 0) mk = k;
                if (!((act >> s) & 1u) && sl[s] < 1.0) {
                    const double num = cc - DtR[k], den = 1.0 - sl[s];
                    if (bden == 0.0 || fabs(num) * bden < fabs(bnum) * den) { bnum = num; bden = den; mk = k; }
                }
            }
        }
        const double mine = bden != 0.0 ? bnum * __drcp_rn(bden) : INFINITY;
        double bt = fabs(mine);
        int bk = mk;
        warp_argmin<true>(bt, bk);
        const double step0 = shfl(mine, bk & 31);  // lane (bk & 31) owns atom bk and offered exactly it
        double step = step0;
        cur = bk;
        double coeff1 = lane <= i ? sg * ul : 0.0, coeff2 = lane <= i ? dl * ul : 0.0;  // two interleaved butterfly sums
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            coeff1 += shfl_xor(coeff1, o);
            coeff2 += shfl_xor(coeff2, o);
        }
        const double step_max2 = cc - lambda1;
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        if (lane <= i) {
            coef_l = fma(step, ul, coef_l);
            if (coef_l < 0.0) coef_l = 0.0;
        }
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            int k = lane + 32 * s;
            if (k < K) DtR[k] = fma(-step, sl[s], DtR[k]);
        }
        normX += coeff1 * step * step - 2.0 * coeff2 * step;
        __syncwarp();
        if (step == step_max) {
            const int z = fz;
            const int az = ind[z];
            const double schur_r = Mi[tri(z, z)];
            double uk = 0.0;
            if (lane < i) uk = (lane < z) ? Mi[tri(z, lane)] : Mi[tri(lane + 1, z)];
            __syncwarp();
            if (lane < i) u[lane] = uk;
            double cn = __shfl_down_sync(FULL, coef_l, 1);
            int in_ = __shfl_down_sync(FULL, ind_l, 1);
            const double rn = __shfl_down_sync(FULL, rs_l, 1);
            const double ksum = warp_sum(lane < i ? uk : 0.0);
            if (lane >= z && lane < i) { coef_l = cn; ind_l = in_; rs_l = rn; }
            if (lane == i) { coef_l = 0.0; ind_l = -1; rs_l = 0.0; }
            if (lane < i) rs_l = rs_l - uk - uk * ksum / schur_r;  // row sums: without column z, then the rank-1 downdate
            if ((az & 31) == lane) act &= ~(1u << (az >> 5));
#pragma unroll 1
            for (int j = z; j < i; ++j) {
                double mv = 0.0;
                if (lane <= j) mv = Mi[tri(j + 1, lane < z ? lane : lane + 1)];
                __syncwarp();
                if (lane <= j) Mi[tri(j, lane)] = mv;
                __syncwarp();
            }
            if (lane <= i) ind[lane] = ind_l;
            __syncwarp();
            if (lane < i) {
                const double ir = uk / schur_r;
#pragma unroll 1
                for (int k = lane; k < i; ++k) Mi[tri(k, lane)] = fma(-ir, u[k], Mi[tri(k, lane)]);
            }
            __syncwarp();
            newAtom = 0;
            na = i;
            i -= 2;
        } else {
            newAtom = 1;
        }
        if (iter >= length_path - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == L - 1) break;
    }
    {
        const bool on = lane < na && ind_l >= 0 && coef_l > 0.0;
#pragma unroll
        for (int s = 0; s < NPL; ++s) sup[s] = __reduce_or_sync(FULL, (on && (ind_l >> 5) == s) ? (1u << (ind_l & 31)) : 0u);
    }
    if (x) {
        if (lane < na && ind_l >= 0) x[ind_l] = coef_l;
        __syncwarp();
    }
    return overflow;
}

// ------------------------------------------------------------------------------------------------
// warp_nnqp_dense: min 1/2 x'Hx - c'x, x >= 0 for n <= 32 atoms (one per lane) when the minimiser is DENSE -- the
// CylinderZeppelinBall elastic net (amico/models.pyx:615: lambda1 = 0, lambda2 = 4) keeps ~23 of its 26 atoms, so a homotopy
// path that starts from the empty set walks ~25 sequential steps.  This starts from the other end: x0 = W c with the precomputed
// inverse W = H^-1 of the direction (H = G + lambda2 I is well conditioned through the ridge), and when x0 >= 0 it IS the
// minimiser (62 % of the cfg5 voxels).  Otherwise block principal pivoting (Judice & Pires) on the zero set R: with x_R = 0,
//   mu = (W_RR)^-1 x0_R,  x = x0 - W[:, R] mu,  gradient on R = -mu,
// exchange { i in F : x_i < 0 } and { i in R : mu_i > 0 } until none is left -- k = |R| stays small (26 - 23), so each round is a
// k x k solve and k rows of W.  H is strictly convex, the minimiser is unique and equals what the reference's LARS reaches; the
// result is CERTIFIED before it is accepted: the KKT conditions are evaluated against H itself (not W) and a voxel that fails them
// -- or does not converge in 24 rounds -- returns false and is solved by the homotopy path instead.
// c, H, W: lane j holds c_j; Hd / Wd: the direction's n x n tables (ld ldH / ldW).  wk: >= 32 * 33 doubles of per-warp scratch.
constexpr int NNQP_KMAX = 12;  // largest zero set the block-pivoting rounds handle (scratch: KMAX x (KMAX + 1) doubles)

__device__ __noinline__ bool warp_nnqp_dense(const double *__restrict__ Hd, int ldH, const double *__restrict__ Wd, int ldW, int n, double c,
                                             double *wk, double &x_out, int lane)
{
    constexpr int LDK = NNQP_KMAX + 1;
    const bool in = lane < n;
    // x0 = W c (W symmetric: row k read coalesced), two rows in flight
    double x0 = 0.0;
    #pragma unroll 1
    for (int k = 0; k < n; k += 2) {
        const double w0 = in ? Wd[(size_t)k * ldW + lane] : 0.0;
        const double w1 = (in && k + 1 < n) ? Wd[(size_t)(k + 1) * ldW + lane] : 0.0;
        x0 = fma(w0, shfl(c, k), x0);
        x0 = fma(w1, shfl(c, min(k + 1, n - 1)), x0);
    }
    unsigned R = __ballot_sync(FULL, in && x0 < 0.0);
    double x = x0;
    if (R != 0u) {
        bool converged = false;
        int best = 33, credit = 3;
        #pragma unroll 1
        for (int round = 0; round < 24; ++round) {
            const int k = __popc(R);
            if (k > NNQP_KMAX) return false;
            // augmented k x (k + 1) system [W_RR | x0_R] in scratch (row per lane), Gauss-Jordan without pivoting (W_RR is SPD)
            const int myrow = (lane < k) ? (int)__fns(R, 0, lane + 1) : 0;  // lane i < k owns the i-th atom of R
            const double x0r = shfl(x0, myrow);
            double *row = wk + lane * LDK;
            if (lane < k) {
                #pragma unroll 1
                for (int q = 0; q < k; ++q) row[q] = Wd[(size_t)myrow * ldW + (int)__fns(R, 0, q + 1)];
                row[k] = x0r;
            }
            __syncwarp();
            #pragma unroll 1
            for (int pv = 0; pv < k; ++pv) {
                const double piv = wk[pv * LDK + pv];
                if (lane < k && lane != pv) {
                    const double f = row[pv] / piv;
                    #pragma unroll 1
                    for (int q = pv + 1; q <= k; ++q) row[q] = fma(-f, wk[pv * LDK + q], row[q]);
                }
                __syncwarp();
            }
            double mu_i = 0.0;  // lane i < k: multiplier of its atom
            if (lane < k) mu_i = row[k] / row[lane];
            __syncwarp();
            // x = x0 - W[:, R] mu; multiplier by atom
            double mu = 0.0;
            x = x0;
            #pragma unroll 1
            for (int i = 0; i < k; ++i) {
                const int r = (int)__fns(R, 0, i + 1);
                const double m_i = shfl(mu_i, i);
                if (in) x = fma(-Wd[(size_t)r * ldW + lane], m_i, x);
                if (lane == r) mu = m_i;
            }
            const bool inR = (R >> lane) & 1u;
            if (inR) x = 0.0;
            const unsigned V = __ballot_sync(FULL, in && ((!inR && x < 0.0) || (inR && mu > 0.0)));
            if (V == 0u) { converged = true; break; }
            const int ninf = __popc(V);
            if (ninf < best) { best = ninf; credit = 3; R ^= V; }
            else if (credit > 0) { --credit; R ^= V; }
            else R ^= 1u << (31 - __clz(V));  // backup rule: a single exchange, highest index
        }
        if (!converged) return false;
    }
    // certificate: g = H x - c;  g_j = 0 (relative) where x_j > 0, g_j >= 0 where x_j = 0
    double g = -c, scale = fabs(c);
    #pragma unroll 1
    for (int k = 0; k < n; k += 2) {
        const double h0 = in ? Hd[(size_t)k * ldH + lane] : 0.0;
        const double h1 = (in && k + 1 < n) ? Hd[(size_t)(k + 1) * ldH + lane] : 0.0;
        g = fma(h0, shfl(x, k), g);
        g = fma(h1, shfl(x, min(k + 1, n - 1)), g);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) scale = fmax(scale, shfl_xor(scale, o));
    const double tol = 1e-10 * scale;
    const bool bad = in && ((x > 0.0) ? (fabs(g) > tol) : (x < 0.0 || g < -tol));
    if (__any_sync(FULL, bad)) return false;
    x_out = in ? x : 0.0;
    return true;
}

}  // namespace amx
