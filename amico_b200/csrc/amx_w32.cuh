// NODDI stage kernels, "group" organisation: one warp owns a GROUP of G voxels of the same LUT direction.
//
// The active sets of the NODDI solves are tiny (stage 1: ~4 passive atoms, stage 2/3: ~10 of 145), so in the
// warp-per-voxel kernels (amx_kernels.cuh::k_noddi_stage) ~90 % of the issued instructions were warp-wide bookkeeping
// around a handful of useful lanes (ncu, profiles/): shuffles, REDUX arg-reductions, predicated triangular solves.
// Here the work is split by shape instead:
//   phase A (warp-cooperative, one voxel after the other): the only O(n * |P|) part -- the dual / correlation pass over all
//           atoms, lane l owning atoms l, l+32, ... -- reads the rows of the per-direction Gram table from SHARED memory
//           (the whole 145x148 fp64 table of the tile's direction is staged once per tile by one TMA bulk copy) and ends
//           in one warp arg-reduction that hands the entering candidate to the voxel's lane;
//   phase B (one LANE per voxel, all voxels of the group at once): everything that is O(|P|^2) -- Cholesky append,
//           triangular solves, ratio test, Givens downdate (NNLS); inverse update / downdate, path direction, step
//           (LARS) -- as plain sequential code on thread-private arrays, no cross-lane traffic at all.
// Arithmetic and pivoting rules are those of warp_nnls / warp_lars_fast (amx_solvers.cuh), i.e. of the reference's
// nnls / lasso calls (amico/models.pyx:911, 926, 940); voxels whose active set outgrows CAP go to the scalar slow path.
#pragma once
#include "amx_kernels.cuh"

namespace amx {

template <int STAGE> struct W32Cfg;
template <> struct W32Cfg<1> { static constexpr int G = 16, CAP = 10; };
template <> struct W32Cfg<2> { static constexpr int G = 8, CAP = 24; };
template <> struct W32Cfg<3> { static constexpr int G = 8, CAP = 25; };

// bytes of per-warp solver state in shared memory; every per-voxel array is laid out [index][G] (voxel-minor), so that the
// lane-per-voxel phase is conflict-free and the warp-cooperative phase reads broadcasts
__host__ __device__ inline unsigned w32_state_bytes(int stage, int npl)
{
    const int G = stage == 1 ? W32Cfg<1>::G : stage == 2 ? W32Cfg<2>::G : W32Cfg<3>::G;
    const int CAP = stage == 1 ? W32Cfg<1>::CAP : stage == 2 ? W32Cfg<2>::CAP : W32Cfg<3>::CAP;
    unsigned nd = (unsigned)(CAP * (CAP + 1) / 2 + 4 * CAP);  // factor / inverse + four CAP-vectors
    if (stage == 2) nd += 1;                                  // ||y2||^2
    unsigned b = nd * G * 8;
    b += 2u * npl * G * 4;                                    // excl, base
    b += (unsigned)(CAP * G);                                 // atom ids (u8)
    return (b + 15u) & ~15u;
}

// packed lower-triangular symmetric table: element (a, b) at tri(max, min).  Consecutive lanes reading T[pk][j] hit
// consecutive words for j <= pk and words tri(j) + pk for j > pk -- triangular numbers are a permutation mod 16, so both
// patterns are bank-conflict free for 64-bit accesses.
__host__ __device__ inline size_t w32_packed_stride(int K) { return ((size_t)K * (K + 1) / 2 + 1) & ~(size_t)1; }

__global__ void k_pack_sym(const double *__restrict__ T, int K, int ld, size_t stride, double *Tp, size_t pstride)
{
    const double *Td = T + (size_t)blockIdx.x * stride;
    double *Pd = Tp + (size_t)blockIdx.x * pstride;
    #pragma unroll 1
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
        const int r = e / K, c = e - r * K;
        if (c <= r) Pd[tri(r, c)] = Td[(size_t)r * ld + c];
    }
}

__device__ __forceinline__ double sym_ld(const double *sT, int a, int b) { return sT[a >= b ? tri(a, b) : tri(b, a)]; }

// ------------------------------------------------------------------------------------------------
// Cholesky downdate: passive position q leaves a factor of pn rows (Givens rotations, z = L^-1 c_P rotated along);
// the sequential twin of the downdate in warp_nnls.
// All arrays are per-voxel shared-memory arrays with element stride G (pointer already offset by the voxel).
template <int G>
__device__ __forceinline__ void lane_downdate(double *L, double *rd, double *z, double *e, int pn, int q)
{
    #pragma unroll 1
    for (int i = q; i < pn - 1; ++i) {
        #pragma unroll 4
        for (int col = 0; col <= i; ++col) L[tri(i, col) * G] = L[tri(i + 1, col) * G];
        e[i * G] = L[tri(i + 1, i + 1) * G];
    }
    #pragma unroll 1
    for (int r = q; r < pn - 1; ++r) {
        const double a = L[tri(r, r) * G], b = e[r * G];
        const double ir = 1.0 / sqrt(fma(a, a, b * b));
        const double cs = a * ir, sn = b * ir;
        const double d = fma(cs, a, sn * b);
        L[tri(r, r) * G] = d;
        rd[r * G] = 1.0 / d;
        #pragma unroll 2
        for (int i = r + 1; i < pn - 1; ++i) {
            const double u1 = L[tri(i, r) * G], u2 = L[tri(i, r + 1) * G];
            L[tri(i, r) * G] = fma(cs, u1, sn * u2);
            L[tri(i, r + 1) * G] = fma(cs, u2, -sn * u1);
        }
        const double zr = z[r * G], zr1 = z[(r + 1) * G];
        z[r * G] = fma(cs, zr, sn * zr1);
        z[(r + 1) * G] = fma(cs, zr1, -sn * zr);
    }
    z[(pn - 1) * G] = 0.0;
}

// bits of word s that do NOT name an atom < n
__device__ __forceinline__ unsigned beyond_mask(int n, int s)
{
    const int r = n - 32 * s;
    return r >= 32 ? 0u : (r <= 0 ? 0xffffffffu : (0xffffffffu << r));
}

// ------------------------------------------------------------------------------------------------
// Stages 1 and 3: Lawson-Hanson NNLS in Gram space for the G voxels of a group.
//   sT  : packed Gram table of the direction in shared memory, scr: c = A^T y of the group's voxels [G][NA] (global)
template <int STAGE, int NPL>
__device__ __noinline__ void w32_nnls_group(const FitParams &p, const double *sT, const double *scr, int nvox, long long pos0,
                                            unsigned char *wst, int lane)
{
    constexpr int G = W32Cfg<STAGE>::G, CAP = W32Cfg<STAGE>::CAP, TRIC = CAP * (CAP + 1) / 2;
    const int v = lane & (G - 1);
    double *Lb = (double *)wst;                          // [TRIC][G] Cholesky factor of H_PP, packed rows
    double *rdb = Lb + TRIC * G;                         // [CAP][G] reciprocal diagonal
    double *zb = rdb + CAP * G;                          // [CAP][G] z = L^-1 c_P
    double *svb = zb + CAP * G;                          // [CAP][G] passive solution / downdate scratch
    double *xs = svb + CAP * G;                          // [CAP][G] coefficients of the passive positions
    unsigned *excl = (unsigned *)(xs + CAP * G);         // [NPL][G] atoms phase A must skip (not allowed | passive | rejected)
    unsigned *base = excl + NPL * G;                     // [NPL][G] not allowed
    unsigned char *Ps = (unsigned char *)(base + NPL * G);  // [CAP][G] passive atoms
    double *L = Lb + v, *rd = rdb + v, *z = zb + v, *sv = svb + v;
    const int n = p.n, NA = p.NA, mcap = p.m;
    int np = 0, iter = 0, nrej = 0, cand = -1, ovf = 0, support = 0;
    bool done = !(lane < nvox);
    int tj[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) tj[s] = tri(lane + 32 * s, 0);
    if (!done) {
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            unsigned b = beyond_mask(n, s);
            if (STAGE == 3) b |= ~p.supmask[(size_t)(pos0 + v) * NPL + s];
            base[s * G + v] = b;
            excl[s * G + v] = b;
            support += __popc(~b);
        }
    }
    const int itmax = STAGE == 3 ? 3 * support : 3 * n;
    __syncwarp();
    unsigned active = __ballot_sync(FULL, !done);
    while (active) {
        // ---- phase A: dual w = c - T[:,P] x_P and its arg-max, voxel after voxel
        unsigned todo = active;
        #pragma unroll 1
        while (todo) {
            const int vv = __ffs(todo) - 1;
            todo &= todo - 1;
            const int npv = __shfl_sync(FULL, np, vv);
            const double *cv = scr + (size_t)vv * NA + lane;
            double w[NPL];
#pragma unroll
            for (int s = 0; s < NPL; ++s) w[s] = __ldcg(cv + 32 * s);
            #pragma unroll 2
            for (int k = 0; k < npv; ++k) {
                const int pk = Ps[k * G + vv];
                const double xk = xs[k * G + vv];
                const int tpk = tri(pk, 0) + lane;
#pragma unroll
                for (int s = 0; s < NPL; ++s) {
                    const int idx = (lane + 32 * s <= pk) ? tpk + 32 * s : tj[s] + pk;
                    w[s] = fma(-sT[idx], xk, w[s]);
                }
            }
            double bv = 0.0;
            int bj = -1;
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const unsigned e = excl[s * G + vv];
                if (!((e >> lane) & 1u) && w[s] > bv) { bv = w[s]; bj = lane + 32 * s; }
            }
            warp_argmax(bv, bj);
            if (lane == vv) cand = (bj >= 0 && bv > 0.0) ? bj : -1;
        }
        __syncwarp();
        // ---- phase B: one lane per voxel
        if (!done) {
            if (cand < 0) {
                done = true;
            } else if (np >= CAP) {
                done = true;
                ovf = 1;
            } else {
                const int j = cand;
                double vq = 0.0, vz = 0.0;
                #pragma unroll 1
                for (int a = 0; a < np; ++a) {  // forward substitution: new factor row = L^-1 T[P, j]
                    double t = sym_ld(sT, (int)Ps[a * G + v], j);
                    #pragma unroll 4
                    for (int i = 0; i < a; ++i) t = fma(-L[tri(a, i) * G], L[tri(np, i) * G], t);
                    const double va = t * rd[a * G];
                    L[tri(np, a) * G] = va;
                    vq = fma(va, va, vq);
                    vz = fma(va, z[a * G], vz);
                }
                const double d2 = sT[tri(j, j)] - vq;
                bool ok = false;
                double znew = 0.0, dd = 0.0;
                if (d2 > 0.0) {
                    const double unorm = sqrt(vq);
                    dd = sqrt(d2);
                    const double tt = unorm + dd * 0.01;
                    if (tt - unorm > 0.0) {
                        znew = (__ldcg(scr + (size_t)v * NA + j) - vz) / dd;
                        ok = znew > 0.0;
                    }
                }
                if (!ok) {  // reject: phase A offers the next best candidate
                    excl[(j >> 5) * G + v] |= 1u << (j & 31);
                    ++nrej;
                } else {
                    L[tri(np, np) * G] = dd;
                    rd[np * G] = 1.0 / dd;
                    z[np * G] = znew;
                    Ps[np * G + v] = (unsigned char)j;
                    xs[np * G + v] = 0.0;
                    ++np;
                    if (nrej) {  // a new outer iteration forgets the rejected candidates
#pragma unroll
                        for (int s = 0; s < NPL; ++s) excl[s * G + v] = base[s * G + v];
                        #pragma unroll 1
                        for (int k = 0; k < np; ++k) {
                            const int a = Ps[k * G + v];
                            excl[(a >> 5) * G + v] |= 1u << (a & 31);
                        }
                        nrej = 0;
                    } else {
                        excl[(j >> 5) * G + v] |= 1u << (j & 31);
                    }
                    bool capped = false;
                    for (;;) {  // secondary loop
                        if (++iter > itmax) { capped = true; break; }
                        #pragma unroll 1
                        for (int i = np - 1; i >= 0; --i) {  // back substitution s = L^-T z
                            double t = z[i * G];
                            #pragma unroll 4
                            for (int k = np - 1; k > i; --k) t = fma(-L[tri(k, i) * G], sv[k * G], t);
                            sv[i * G] = t * rd[i * G];
                        }
                        bool anyneg = false;
                        int cnd = -1;
                        double tmin = INFINITY;
                        #pragma unroll 1
                        for (int k = 0; k < np; ++k) {
                            const double sk = sv[k * G];
                            if (sk <= 0.0) {
                                anyneg = true;
                                const double xp = xs[k * G + v];
                                const double tt = -xp / (sk - xp);
                                if (tt < 2.0 && (cnd < 0 || tt < tmin)) { tmin = tt; cnd = k; }
                            }
                        }
                        if (!anyneg || cnd < 0) break;
                        unsigned removed = 0;
                        int wr = 0;
                        #pragma unroll 1
                        for (int k = 0; k < np; ++k) {
                            double xp = xs[k * G + v];
                            xp = fma(tmin, sv[k * G] - xp, xp);
                            if (k == cnd) xp = 0.0;
                            const int a = Ps[k * G + v];
                            if (xp > 0.0) {
                                Ps[wr * G + v] = (unsigned char)a;
                                xs[wr * G + v] = xp;
                                ++wr;
                            } else {
                                removed |= 1u << k;
                                excl[(a >> 5) * G + v] &= ~(1u << (a & 31));
                            }
                        }
                        int pn = np;
                        np = wr;
                        if (np == 0) break;
                        while (removed) {  // highest position first
                            const int q = 31 - __clz(removed);
                            removed &= ~(1u << q);
                            lane_downdate<G>(L, rd, z, sv, pn, q);
                            --pn;
                        }
                    }
                    if (capped) {
                        done = true;
                    } else {
                        #pragma unroll 1
                        for (int k = 0; k < np; ++k) xs[k * G + v] = sv[k * G];
                        if (np >= mcap) done = true;
                    }
                }
            }
            if (done) {  // this voxel is finished: results
                const long long pos = pos0 + v;
                const long long vox = (long long)p.order[pos];
                if (ovf) {
                    unsigned long long idx = atomicAdd((unsigned long long *)&p.status[2], 1ull);
                    if ((long long)idx < p.ovf_cap) p.ovf_list[idx] = (int)vox;
                }
                if (STAGE == 1) {
                    double xi = 0.0, xd = 0.0;
                    #pragma unroll 1
                    for (int k = 0; k < np; ++k) {
                        const int a = Ps[k * G + v];
                        if (a == n - 1) xi = xs[k * G + v];
                        if (p.exvivo && a == n - 2) xd = xs[k * G + v];
                    }
                    if (ovf) xi = xd = 0.0;
                    p.xiso[2 * pos] = xi;
                    p.xiso[2 * pos + 1] = xd;
                } else if (!ovf) {
                    // maps (amico/models.pyx:945-979): sums run over the positive coefficients in atom order
                    #pragma unroll 1
                    for (int k = 1; k < np; ++k) {  // insertion sort of the passive set by atom index
                        const int a = Ps[k * G + v];
                        const double xa = xs[k * G + v];
                        int q = k - 1;
                        while (q >= 0 && (int)Ps[q * G + v] > a) {
                            Ps[(q + 1) * G + v] = Ps[q * G + v];
                            xs[(q + 1) * G + v] = xs[q * G + v];
                            --q;
                        }
                        Ps[(q + 1) * G + v] = (unsigned char)a;
                        xs[(q + 1) * G + v] = xa;
                    }
                    const int n_wm = p.n_wm;
                    double s_all = 0.0, s_wm = 0.0, f1 = 0.0, f2 = 0.0, k1 = 0.0, x_iso = 0.0, x_dot = 0.0;
                    #pragma unroll 1
                    for (int k = 0; k < np; ++k) {
                        const double xk = xs[k * G + v];
                        const int a = Ps[k * G + v];
                        if (a == n - 1) x_iso = xk;
                        if (p.exvivo && a == n - 2) x_dot = xk;
                        if (xk > 0.0) s_all += xk;
                    }
                    s_all += 1e-16;
                    #pragma unroll 1
                    for (int k = 0; k < np; ++k) {
                        const double xk = xs[k * G + v];
                        if (xk > 0.0 && (int)Ps[k * G + v] < n_wm) s_wm += xk / s_all;
                    }
                    s_wm += 1e-16;
                    #pragma unroll 1
                    for (int k = 0; k < np; ++k) {
                        const double xk = xs[k * G + v];
                        const int a = Ps[k * G + v];
                        if (xk > 0.0 && a < n_wm) {
                            const float ic = p.icvf[a];
                            f1 += (double)ic * xk / s_all / s_wm;
                            f2 += (double)((float)(1.0 - (double)ic)) * xk / s_all / s_wm;
                            k1 += (double)p.kappa[a] * xk / s_all / s_wm;
                        }
                    }
                    const double ndi = f1 / (f1 + f2 + 1e-16);
                    const double odi = 2.0 / 3.14159265358979323846 * atan2(1.0, k1);
                    const double fwf = x_iso / s_all;
                    double *e = p.est + vox * p.n_maps;
                    e[0] = ndi; e[1] = odi; e[2] = fwf;
                    if (p.exvivo) e[3] = x_dot / s_all;
                    if (p.flags & FLAG_EXTRA) {
                        const double tf = 1.0 - fwf;
                        p.extra[2 * vox] = ndi * tf;
                        p.extra[2 * vox + 1] = odi * tf;
                    }
                    if (p.support_out) p.support_out[vox] = support;
                }
            }
        }
        __syncwarp();
        active = __ballot_sync(FULL, !done);
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 2: the non-negative LARS / homotopy path of warp_lars_fast for the G voxels of a group; only the support of the
// result is consumed (amico/models.pyx:929-936) and goes to supmask.
//   sT: packed T2 = G + ridge I of the direction in shared memory; scr: c2 = A2^T y2 of the group's voxels [G][NA]
template <int NPL>
__device__ __noinline__ void w32_lars_group(const FitParams &p, const double *sT, const double *scr, int nvox, long long pos0,
                                            unsigned char *wst, int lane)
{
    constexpr int G = W32Cfg<2>::G, CAP = W32Cfg<2>::CAP, TRIC = CAP * (CAP + 1) / 2;
    const int v = lane & (G - 1);
    double *Mib = (double *)wst;                  // [TRIC][G] inverse of G_SS, packed
    double *uub = Mib + TRIC * G;                 // [CAP][G] scratch
    double *gsb = uub + CAP * G;                  // [CAP][G] scratch
    double *xs = gsb + CAP * G;                   // [CAP][G] coefficients of the active positions
    double *us = xs + CAP * G;                    // [CAP][G] path direction
    double *nrm = us + CAP * G;                   // [G] ||y2||^2 (written by gemm_c2)
    unsigned *excl = (unsigned *)(nrm + G);       // [NPL][G] active atoms | beyond K
    unsigned *base = excl + NPL * G;              // [NPL][G] beyond K
    unsigned char *Ss = (unsigned char *)(base + NPL * G);  // [CAP][G] active atoms
    double *Mi = Mib + v, *uu = uub + v, *gs = gsb + v;
    const int K = p.n_wm, NA = p.NA, n = p.n;
    const int Lmax = p.dc < K ? p.dc : K;
    const double lambda1 = p.lambda1;
    bool done = !(lane < nvox);
    int i = 0, iter = 0, newAtom = 1, cur = -1, na = 0, fz = -1, ovf = 0;
    double normX = done ? 0.0 : nrm[v], step_max = INFINITY, cstep = 0.0, ccv = 0.0;
    int tj[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) tj[s] = tri(lane + 32 * s, 0);
    if (!done) {
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            const unsigned b = beyond_mask(K, s);
            base[s * G + v] = b;
            excl[s * G + v] = b;
        }
    }
    __syncwarp();
    // ---- round 0: most correlated atom (largest DtR, lowest index)
    {
        unsigned todo = __ballot_sync(FULL, !done);
        #pragma unroll 1
        while (todo) {
            const int vv = __ffs(todo) - 1;
            todo &= todo - 1;
            const double *cv = scr + (size_t)vv * NA + lane;
            double bv = 0.0;
            int bi = -1;
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const int k = lane + 32 * s;
                if (k < K) {
                    const double d = __ldcg(cv + 32 * s);
                    if (bi < 0 || d > bv) { bv = d; bi = k; }
                }
            }
            warp_argmax(bv, bi);
            if (lane == vv) { cur = bi; ccv = bv; }
        }
        if (!done && (Lmax <= 0 || cur < 0 || fabs(ccv) < lambda1)) done = true;  // empty support
    }
    bool first = true;
    unsigned active = __ballot_sync(FULL, true);  // enter the loop once so that finished voxels still write their mask
    while (active) {
        // ---- phase B: finish the previous step, start the next one (add / drop an atom, new path direction)
        bool fin = false;
        if (!done) {
            if (!first) {
                // step length (warp_lars_fast): candidate from phase A, lambda reached, or an active coefficient hitting zero
                double coeff1 = 0.0;
                #pragma unroll 1
                for (int l = 0; l <= i; ++l) coeff1 += us[l * G + v];
                const double cc = fabs(ccv);
                const double coeff2 = cc * coeff1;
                const double step_max2 = cc - lambda1;
                const double step = fmin(fmin(cstep, step_max2), step_max);
                if (step == INFINITY) {
                    fin = true;
                } else {
                    #pragma unroll 1
                    for (int l = 0; l <= i; ++l) {
                        double c = fma(step, us[l * G + v], xs[l * G + v]);
                        xs[l * G + v] = c < 0.0 ? 0.0 : c;
                    }
                    normX += coeff1 * step * step - 2.0 * coeff2 * step;
                    if (step == step_max) {  // drop active position fz: shrink the lists, downdate the inverse
                        const int zq = fz;
                        const int az = Ss[zq * G + v];
                        const double schur_r = Mi[tri(zq, zq) * G];
                        #pragma unroll 1
                        for (int l = 0; l < i; ++l) uu[l * G] = (l < zq) ? Mi[tri(zq, l) * G] : Mi[tri(l + 1, zq) * G];
                        #pragma unroll 1
                        for (int l = zq; l < i; ++l) {
                            Ss[l * G + v] = Ss[(l + 1) * G + v];
                            xs[l * G + v] = xs[(l + 1) * G + v];
                        }
                        excl[(az >> 5) * G + v] &= ~(1u << (az & 31));
                        #pragma unroll 1
                        for (int jc = zq; jc < i; ++jc) {
                            #pragma unroll 1
                            for (int l = 0; l <= jc; ++l) Mi[tri(jc, l) * G] = Mi[tri(jc + 1, l < zq ? l : l + 1) * G];
                        }
                        #pragma unroll 1
                        for (int l = 0; l < i; ++l) {
                            const double ir = uu[l * G] / schur_r;
                            #pragma unroll 1
                            for (int k = l; k < i; ++k) Mi[tri(k, l) * G] = fma(-ir, uu[k * G], Mi[tri(k, l) * G]);
                        }
                        newAtom = 0;
                        na = i;
                        i -= 2;
                    } else {
                        newAtom = 1;
                    }
                    if (iter >= 4 * Lmax - 1 || fabs(step) < 1e-15 || step == step_max2 || normX < 1e-15 || i == Lmax - 1) fin = true;
                }
                if (!fin) {
                    ++i;
                    if (i >= Lmax || i < 0) fin = true;
                }
            }
            if (!fin) {
                ++iter;
                if (newAtom) {
                    if (i >= CAP) {
                        ovf = 1; na = i; fin = true;
                    } else {
                        Ss[i * G + v] = (unsigned char)cur;
                        xs[i * G + v] = 0.0;
                        excl[(cur >> 5) * G + v] |= 1u << (cur & 31);
                        #pragma unroll 1
                        for (int l = 0; l <= i; ++l) gs[l * G] = sym_ld(sT, cur, (int)Ss[l * G + v]);
                        if (i == 0) {
                            Mi[0] = 1.0 / gs[0];
                        } else {
                            double dot = 0.0;
                            #pragma unroll 1
                            for (int r = 0; r < i; ++r) {
                                double ur = 0.0;
                                #pragma unroll 1
                                for (int c = 0; c < i; ++c) ur = fma(Mi[(r <= c ? tri(c, r) : tri(r, c)) * G], gs[c * G], ur);
                                uu[r * G] = ur;
                                dot = fma(ur, gs[r * G], dot);
                            }
                            const double schur = 1.0 / (gs[i * G] - dot);
                            #pragma unroll 1
                            for (int r = 0; r < i; ++r) {
                                const double su = schur * uu[r * G];
                                #pragma unroll 1
                                for (int k = r; k < i; ++k) Mi[tri(k, r) * G] = fma(su, uu[k * G], Mi[tri(k, r) * G]);
                                Mi[tri(i, r) * G] = -su;
                            }
                            Mi[tri(i, i) * G] = schur;
                        }
                    }
                }
                if (!fin) {
                    na = i + 1;
                    // path direction u = (G_SS)^-1 1 and the largest step before an active coefficient crosses zero
                    step_max = INFINITY;
                    fz = -1;
                    #pragma unroll 1
                    for (int l = 0; l <= i; ++l) {
                        double ul = 0.0;
                        #pragma unroll 1
                        for (int c = 0; c <= i; ++c) ul += Mi[(l <= c ? tri(c, l) : tri(l, c)) * G];
                        us[l * G + v] = ul;
                        const double r = -xs[l * G + v] / ul;
                        if (r > 0.0 && r <= step_max) { step_max = r; fz = l; }
                    }
                }
            }
            if (fin) done = true;
        }
        if (lane < nvox && done && (fin || first)) {  // support mask (amico/models.pyx:929-936), once per voxel
            const long long pos = pos0 + v;
            unsigned wds[NPL];
#pragma unroll
            for (int s = 0; s < NPL; ++s) wds[s] = ~beyond_mask(n, s) & beyond_mask(K, s);  // dot / iso columns
            if (!ovf) {
                #pragma unroll 1
                for (int l = 0; l < na; ++l) {
                    const int a = Ss[l * G + v];
                    if (xs[l * G + v] > 0.0) {
#pragma unroll
                        for (int s = 0; s < NPL; ++s)
                            if (s == (a >> 5)) wds[s] |= 1u << (a & 31);
                    }
                }
            } else {
                unsigned long long idx = atomicAdd((unsigned long long *)&p.status[2], 1ull);
                if ((long long)idx < p.ovf_cap) p.ovf_list[idx] = p.order[pos];
            }
#pragma unroll
            for (int s = 0; s < NPL; ++s) p.supmask[(size_t)pos * NPL + s] = wds[s];
        }
        first = false;
        __syncwarp();
        active = __ballot_sync(FULL, !done);
        // ---- phase A: correlations of all atoms with the residual and their slopes along u; entering candidate
        unsigned todo = active;
        #pragma unroll 1
        while (todo) {
            const int vv = __ffs(todo) - 1;
            todo &= todo - 1;
            const int nav = __shfl_sync(FULL, na, vv);
            const double *cv = scr + (size_t)vv * NA + lane;
            double a[NPL], b[NPL];
#pragma unroll
            for (int s = 0; s < NPL; ++s) { a[s] = __ldcg(cv + 32 * s); b[s] = 0.0; }
            #pragma unroll 2
            for (int l = 0; l < nav; ++l) {
                const int sl = Ss[l * G + vv];
                const double xl = xs[l * G + vv], ul = us[l * G + vv];
                const int tsl = tri(sl, 0) + lane;
#pragma unroll
                for (int s = 0; s < NPL; ++s) {
                    const double r = sT[(lane + 32 * s <= sl) ? tsl + 32 * s : tj[s] + sl];
                    a[s] = fma(-r, xl, a[s]);  // DtR = c2 - T[:,S] x
                    b[s] = fma(r, ul, b[s]);   // slope = T[:,S] u
                }
            }
            // common correlation = DtR of the first active atom
            const int s0 = Ss[vv];
            double d0 = 0.0;
#pragma unroll
            for (int s = 0; s < NPL; ++s)
                if (s == (s0 >> 5)) d0 = a[s];
            const double cc = fabs(shfl(d0, s0 & 31));
            double bt = INFINITY, tb = 0.0;
            int bk = -1;
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const int k = lane + 32 * s;
                const unsigned e = excl[s * G + vv], eb = base[s * G + vv];
                if (!((eb >> lane) & 1u)) {
                    double tl = INFINITY;
                    if (!((e >> lane) & 1u) && b[s] < 1.0) tl = (cc - a[s]) * __drcp_rn(1.0 - b[s]);
                    const double at = fabs(tl);
                    if (bk < 0 || at < bt) { bt = at; bk = k; tb = tl; }
                }
            }
            warp_argmin<true>(bt, bk);
            const double stp = shfl(tb, bk & 31);  // lane (bk & 31) owns atom bk, and its local best is the winner
            if (lane == vv) { cur = bk; cstep = stp; ccv = cc; }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// One CTA per SM; tiles = up to TILE voxels of one direction (global queue).  Per tile one thread issues a single TMA bulk
// copy of the direction's Gram table into shared memory; warps pull groups of G voxels from the tile, run the A^T Y
// micro-GEMMs (DMMA, as in k_noddi_stage -- they do not need the table, so they overlap the copy) and solve the group.
template <int STAGE, int NPL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_noddi_w32(const FitParams p)
{
    constexpr int G = W32Cfg<STAGE>::G;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *mbar = (uint64_t *)smem;
    int *s_tile = (int *)(smem + 8);
    int *s_next = (int *)(smem + 12);
    const double *sT = (const double *)(smem + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wst = smem + 128 + p.w32_T_bytes[STAGE - 1] + (size_t)warp * p.w32_state[STAGE - 1];
    constexpr int NT = 4 * NPL, TP = NT;
    const int m = p.m, n = p.n, n_pad = p.n_pad, n_wm = p.n_wm, NA = p.NA;
    double *scr = p.scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * (size_t)32 * NA;
    const int g = lane >> 2;
    int *counter = p.tile_counter + (STAGE - 1);
    const int n_tiles = *p.n_tiles_ptr;
    const double *Tg = STAGE == 2 ? p.T2p : p.T1p;  // packed symmetric tables
    const size_t T_stride = STAGE == 2 ? p.T2p_stride : p.T1p_stride;
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        mbar_fence_init();
    }
    for (;;) {
        __syncthreads();  // every warp is done with the previous tile's table
        if (threadIdx.x == 0) {
            const int t = atomicAdd(counter, 1);
            *s_tile = t;
            *s_next = 0;
            if (t < n_tiles) {
                const int dir = p.tiles[t].x;
                mbar_expect_tx(mbar, p.w32_T_bytes[STAGE - 1]);
                bulk_g2s((void *)sT, Tg + (size_t)dir * T_stride, p.w32_T_bytes[STAGE - 1], mbar);
            }
        }
        __syncthreads();
        const int t = *s_tile;
        if (t >= n_tiles) break;
        const int4 tile = p.tiles[t];
        const float *S = (const float *)p.slab + (size_t)tile.x * p.slab_stride;
        bool ready = false;
        for (;;) {
            int sub = 0;
            if (lane == 0) sub = atomicAdd(s_next, 1);
            sub = __shfl_sync(FULL, sub, 0);
            if (sub * G >= tile.z) break;
            const int nvox = min(G, tile.z - sub * G);
            const long long pos0 = (long long)tile.y + (long long)sub * G;
            #pragma unroll 1
            for (int b0 = 0; b0 < nvox; b0 += BV) {
                const bool vvalid = b0 + g < nvox;
                const long long mypos = pos0 + (vvalid ? b0 + g : 0);
                const long long myvox = (long long)p.order[mypos];
                if (STAGE == 2) {
                    double *nrm = (double *)wst + (W32Cfg<2>::CAP * (W32Cfg<2>::CAP + 1) / 2 + 4 * W32Cfg<2>::CAP) * W32Cfg<2>::G + b0;
                    const double xi = p.xiso[2 * mypos], xd = p.xiso[2 * mypos + 1];
                    if (p.norms_const)
                        gemm_c2<NT, TP, float, true>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo,
                                                     p.norms, n_wm, scr + (size_t)b0 * NA, NA, nrm, lane);
                    else
                        gemm_c2<NT, TP, float, false>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo,
                                                      p.norms, n_wm, scr + (size_t)b0 * NA, NA, nrm, lane);
                } else {
                    gemm_c1<NT, TP, float>(S, n_pad, m, p.y, p.y_f64, myvox, vvalid, scr + (size_t)b0 * NA, NA, lane);
                }
            }
            if (!ready) {
                mbar_wait(mbar, phase);
                ready = true;
            }
            if (STAGE == 2) w32_lars_group<NPL>(p, sT, scr, nvox, pos0, wst, lane);
            else w32_nnls_group<STAGE, NPL>(p, sT, scr, nvox, pos0, wst, lane);
            __syncwarp();
        }
        if (!ready) mbar_wait(mbar, phase);  // the copy must have landed before the table is overwritten
        phase ^= 1;
    }
}

}  // namespace amx
