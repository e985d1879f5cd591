// NODDI stage kernels, second generation: the same solvers as amx_solvers.cuh (same pivoting rules, same floating-point
// operations in the same order -- the maps are bit-identical to the first-generation stage kernels) re-written around the
// instruction count.  ncu on the first generation (profiles/ncu_full_r02_noddi_stage_kernels.json, tools/ncu_sass_annot.py):
// issue-bound at 65-69 % issue-active with ~9.9 k warp instructions per voxel in stage 1, of which only 10 % were fp64 math --
// the rest was glue: 64-bit index arithmetic and descriptor moves (R2UR) around every Gram-row load of the __noinline__
// solver, XOR register swaps after every 64-bit shuffle, spill reloads (LDL) at 64 registers, __fns() loops in the passive-set
// compaction, convergence checks (BRA.DIV) before every shuffle.  Here: the solver is inlined into a dedicated kernel,
// c lives in registers, x lives by passive position (shared-memory broadcast for the dual pass), all offsets are 32-bit, the
// per-warp workspace is addressed from one base with compile-time offsets, compaction goes through shared memory.
#pragma once
#include "amx_kernels.cuh"

namespace amx {

// 64-bit shuffle as two 32-bit shuffles with plain moves around them (the header's version packs through volatile asm,
// which made ptxas emit a three-XOR register swap after every use in the substitution loops)
__device__ __forceinline__ double shfl2(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(FULL, lo, src);
    hi = __shfl_sync(FULL, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl2_xor(double v, int m)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(FULL, lo, m);
    hi = __shfl_xor_sync(FULL, hi, m);
    return __hiloint2double(hi, lo);
}

// per-warp workspace of the lean NNLS (doubles): packed factor, x by passive position, passive list, batch ||y||^2
template <int CAP>
struct LeanWS {
    static constexpr int LP = 0;
    static constexpr int XS = CAP * (CAP + 1) / 2;
    static constexpr int PI = XS + CAP;
    static constexpr int BX = PI + CAP / 2;
    static constexpr int CS = BX + BV;  // c, 32 NPL doubles (the kernels add it to SIZE)
    static constexpr int SIZE = CS;
};

// Lawson-Hanson in Gram space on the full dictionary (warp_nnls<NPL, false> with identical arithmetic); returns the overflow
// flag.  Out: zz = ||z||^2 of the final passive system, x_last / x_prev = coefficients of atoms n - 1 and n - 2.
template <int NPL, int CAP>
__device__ __forceinline__ int nnls_lean(const double *__restrict__ T, const int ldT, const int n, const int mcap, const int itmax,
                                         const double *cs, double *Lp, int *P, double *xs, const int lane, const int cap,
                                         const bool use_as, const ASpace as, double &zz, double &x_last, double &x_prev)
{
    int np = 0, iter = 0, overflow = 0;
    unsigned inP = 0u, avail = 0u;
    double xp = 0.0, zl = 0.0, rdl = 0.0;  // coefficient, z and 1 / diagonal of this lane's passive position
    int myP = 0;                           // atom at this lane's passive position (mirror of P[lane])
#pragma unroll
    for (int s = 0; s < NPL; ++s) avail |= (lane + 32 * s < n ? 1u : 0u) << s;
    const double *Tl = T + lane;
    double *myrow = Lp + tri(lane, 0);  // this lane's row of the factor (dereferenced only while lane < np <= CAP)
    for (;;) {
        if (np >= mcap) break;
        if (np >= cap) { overflow = 1; break; }
        // ---- dual w = c - T[:,P] x_P in passive order, two rows in flight
        double wl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) wl[s] = cs[lane + 32 * s];
        {
            int k = 0;
#pragma unroll 1
            for (; k + 2 <= np; k += 2) {
                const double *r0 = Tl + (unsigned)(P[k] * ldT), *r1 = Tl + (unsigned)(P[k + 1] * ldT);
                double g0[NPL], g1[NPL];
#pragma unroll
                for (int s = 0; s < NPL; ++s) g0[s] = r0[32 * s];
#pragma unroll
                for (int s = 0; s < NPL; ++s) g1[s] = r1[32 * s];
                const double x0 = xs[k], x1 = xs[k + 1];
#pragma unroll
                for (int s = 0; s < NPL; ++s) wl[s] = fma(-g0[s], x0, wl[s]);
#pragma unroll
                for (int s = 0; s < NPL; ++s) wl[s] = fma(-g1[s], x1, wl[s]);
            }
            if (k < np) {
                const double *r0 = Tl + (unsigned)(P[k] * ldT);
                const double x0 = xs[k];
#pragma unroll
                for (int s = 0; s < NPL; ++s) wl[s] = fma(-r0[32 * s], x0, wl[s]);
            }
        }
        // ---- candidate selection (largest positive dual first; near-dependent / non-improving candidates are dropped)
        unsigned valid = avail & ~inP;
        int j;
        double v, d2, znum;
        for (;;) {
            double bv = 0.0;
            int bj = -1;
#pragma unroll
            for (int s = 0; s < NPL; ++s)
                if (((valid >> s) & 1u) && wl[s] > bv) { bv = wl[s]; bj = lane + 32 * s; }
            warp_argmax_pos(bv, bj);
            j = bj;
            if (j < 0) break;
            const double *Tj = T + (unsigned)(j * ldT);
            double t = (lane < np) ? Tj[myP] : 0.0;  // = T[P[lane]][j] (the table is exactly symmetric)
            const double hjj = Tj[j], cj = cs[j];
#pragma unroll 1
            for (int k = 0; k < np; ++k) {  // v = L^-1 t (forward substitution, lane a owns row a)
                const double vk = shfl2(t * rdl, k);
                if (lane > k && lane < np) t = fma(-myrow[k], vk, t);
            }
            v = (lane < np) ? t * rdl : 0.0;
            double vv = v * v, vz = v * zl;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vv += shfl2_xor(vv, o);
                vz += shfl2_xor(vz, o);
            }
            d2 = hjj - vv;
            znum = cj - vz;
            if (use_as && np > 0 && d2 < 1e-10 * hjj) {
                // near-dependent candidate: both quantities re-evaluated in A-space (see warp_nnls for the why)
                double beta = v;
                {
                    const double *col = Lp + tri(np - 1, 0) + lane;
#pragma unroll 1
                    for (int k = np - 1; k >= 0; --k) {
                        const double sk = shfl2(beta * rdl, k);
                        if (lane < k) beta = fma(-*col, sk, beta);
                        col -= k;
                    }
                    beta = (lane < np) ? beta * rdl : 0.0;
                }
                const float *Sj = as.S + j;
                double a2 = 0.0, ay = 0.0;
#pragma unroll 1
                for (int i0 = 0; i0 < as.m; i0 += 32) {
                    const int i = i0 + lane;
                    const bool on = i < as.m;
                    const float *Si = as.S + (size_t)(on ? i : 0) * as.n_pad;
                    double r = (double)Sj[(size_t)(on ? i : 0) * as.n_pad];
#pragma unroll 1
                    for (int a = 0; a < np; a += 2) {
                        const int a1 = min(a + 1, np - 1);
                        const float s0 = Si[P[a]], s1 = Si[P[a1]];
                        const double b0 = shfl2(beta, a), b1 = (a + 1 < np) ? shfl2(beta, a1) : 0.0;
                        r = fma(-(double)s0, b0, r);
                        r = fma(-(double)s1, b1, r);
                    }
                    if (on) {
                        const double yi = as.y_f64 ? ((const double *)as.y)[as.vox * as.m + i] : (double)((const float *)as.y)[as.vox * as.m + i];
                        a2 = fma(r, r, a2);
                        ay = fma(r, yi, ay);
                    }
                }
                d2 = warp_sum(a2);
                znum = warp_sum(ay);
                if (d2 < 1e-24 * hjj) d2 = 0.0;
            }
            if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) break;
            if ((j & 31) == lane) valid &= ~(1u << (j >> 5));
        }
        if (j < 0) break;
        // ---- j joins the passive set: one more row of the factor
        {
            const double ird = rsqrt(d2), dd = d2 * ird;
            double *rownew = Lp + tri(np, 0);
            if (lane < np) rownew[lane] = v;
            if (lane == np) {
                rownew[lane] = dd;
                rdl = ird;
                zl = znum * ird;
                xp = 0.0;
                myP = j;
                P[lane] = j;
            }
            if ((j & 31) == lane) inP |= 1u << (j >> 5);
            ++np;
        }
        __syncwarp();
        // ---- secondary loop: solve on the passive set, step back to the feasible boundary while a coefficient is not positive
        double s;
        for (;;) {
            if (++iter > itmax) goto done;
            {   // s = L^-T z (back substitution, lane a owns column a)
                s = (lane < np) ? zl : 0.0;
                const double *col = Lp + tri(np - 1, 0) + lane;  // element (k, lane) of row k, walking up
#pragma unroll 1
                for (int k = np - 1; k >= 0; --k) {
                    const double sk = shfl2(s * rdl, k);
                    if (lane < k) s = fma(-*col, sk, s);
                    col -= k;
                }
                s = (lane < np) ? s * rdl : 0.0;
            }
            const bool neg = (lane < np) && (s <= 0.0);
            if (!__any_sync(FULL, neg)) break;
            double tmin = INFINITY;
            int cand = -1;
            if (neg) {
                const double tt = -xp / (s - xp);
                if (tt < 2.0) { tmin = tt; cand = lane; }
            }
            warp_argmin<true>(tmin, cand);
            if (cand < 0) break;
            if (lane < np) xp = fma(tmin, s - xp, xp);
            if (lane == cand) xp = 0.0;
            const bool keep = (lane < np) && (xp > 0.0);
            const unsigned kmask = __ballot_sync(FULL, keep);
            const int np_old = np;
            unsigned rmask = ~kmask & ((1u << np_old) - 1u);  // removed positions (np_old <= CAP < 32)
            for (unsigned r2 = rmask; r2; r2 &= r2 - 1) {    // their atoms leave the passive bit set
                const int a = __shfl_sync(FULL, myP, __ffs(r2) - 1);
                if ((a & 31) == lane) inP &= ~(1u << (a >> 5));
            }
            // compaction of (atom, coefficient) through shared memory: kept position -> its rank among the kept ones
            np = __popc(kmask);
            if (keep) {
                const int rank = __popc(kmask & ((1u << lane) - 1u));
                P[rank] = myP;
                xs[rank] = xp;
            }
            __syncwarp();
            if (lane < np) { myP = P[lane]; xp = xs[lane]; }
            else xp = 0.0;
            if (np == 0) break;
            // Cholesky downdate (column deletion + Givens), one removed position at a time, highest first
            for (int pn = np_old; rmask; --pn) {
                const int q = 31 - __clz(rmask);
                rmask &= ~(1u << q);
                chol_delete(Lp, q, pn, rdl, zl, lane);
            }
        }
        if (lane < np) { xp = s; xs[lane] = s; }
        __syncwarp();
    }
done:
    {
        const unsigned m1 = __ballot_sync(FULL, lane < np && myP == n - 1), m2 = __ballot_sync(FULL, lane < np && myP == n - 2);
        x_last = m1 ? shfl2(xp, __ffs(m1) - 1) : 0.0;
        x_prev = m2 ? shfl2(xp, __ffs(m2) - 1) : 0.0;
    }
    zz = warp_sum(lane < np ? zl * zl : 0.0);
    return overflow;
}

// ------------------------------------------------------------------------------------------------
// NODDI stage 1 (isotropic fraction, amico/models.pyx:911) on the lean solver; same batch queue, same outputs as
// k_noddi_stage<1>.
template <int NPL, int MAXT, int CAP>
__global__ void __launch_bounds__(MAXT, 1) k_noddi_stage1_lean(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *wsb = (double *)(smem + p.ws_smem_off) + warp * (LeanWS<CAP>::SIZE + 32 * NPL);
    double *Lp = wsb + LeanWS<CAP>::LP, *xs = wsb + LeanWS<CAP>::XS, *bx = wsb + LeanWS<CAP>::BX, *cs = wsb + LeanWS<CAP>::CS;
    int *P = (int *)(wsb + LeanWS<CAP>::PI);
    constexpr int NT = 4 * NPL, TP = (MAXT > 512 && NT % 2 == 0) ? NT / 2 : NT;
    const int n = p.n, NA = p.NA;
    const int cap = min(min(p.cap_stage[0], CAP), c_lc_cap);
    double *scr = p.scratch + ((size_t)blockIdx.x * 32 + warp) * (size_t)BV * NA;
    const int g = lane >> 2;
    int *counter = p.tile_counter;
    const int n_tiles = *p.n_tiles_ptr;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1);
        b = __reduce_add_sync(FULL, b);  // a REDUX result is warp-uniform for the compiler (a shuffle's is not): no convergence checks downstream
        if (b >= n_tiles) break;
        const int4 tile = p.tiles[b];
        const int nb = tile.z;  // <= BV
        const float *S = (const float *)p.slab + (size_t)tile.x * p.slab_stride;
        const bool vvalid = g < nb;
        const long long mypos = tile.y + (vvalid ? g : 0);
        const long long myvox = (long long)p.order[mypos];
        const double *T1 = p.T1 + (size_t)tile.x * p.T1_stride;
        double *c1b = p.c1_all ? p.c1_all + (size_t)tile.y * NA : scr;
        gemm_c1<NT, TP, float>(S, p.n_pad, p.m, p.y, p.y_f64, myvox, vvalid, c1b, NA, lane, bx);
#pragma unroll 1
        for (int v = 0; v < nb; ++v) {
            const long long pos = tile.y + v;
            const ASpace asp{S, p.n_pad, p.m, p.y, p.y_f64, (long long)p.order[pos]};
#pragma unroll
            for (int s = 0; s < NPL; ++s) cs[lane + 32 * s] = c1b[(size_t)v * NA + lane + 32 * s];
            __syncwarp();
            double zz, x1, x2;
            const int ov = nnls_lean<NPL, CAP>(T1, p.ldT1, n, p.m, 3 * n, cs, Lp, P, xs, lane, cap, p.aspace != 0, asp,
                                               zz, x1, x2);
            if (lane == 0) {
                p.xiso[2 * pos] = x1;
                p.xiso[2 * pos + 1] = p.exvivo ? x2 : 0.0;
                const double yy = bx[v];
                if (yy > 0.0 && yy - zz < p.exact_tol * yy) {  // exact-fit voxel: queued for the A-space QR path (see k_noddi_stage)
                    const unsigned long long idx = atomicAdd((unsigned long long *)&p.status[4], 1ull);
                    if ((long long)idx < p.exact_cap) p.exact_list[idx] = (int)p.order[pos];
                }
            }
            if (ov) queue_slow(p, (long long)p.order[pos], lane);
            __syncwarp();
        }
    }
}

// ================================================================================================
// Two voxels per warp.  The passive sets of NODDI stage 1 never hold more than 16 atoms, and ~85 % of the solver's instructions are
// sequential glue (substitution steps, reductions, factor updates, control) that a full warp executes for ONE voxel with
// 16-28 lanes idle.  Here each HALF-warp owns a voxel of the same batch (same direction, same Gram table): lane h = lane & 15
// of a half owns atoms h, h + 16, ... (NPH per lane) and passive position h; shuffles run with width 16, reductions with the
// half's member mask, every loop runs to the larger trip count of the two halves with the shorter one masked.  The O(n |P|)
// dual pass costs the same per voxel (each instruction fetches one row for each of the two voxels); everything else is shared.
// Same pivoting rules and the same arithmetic as warp_nnls, except that the two short sums of the candidate test (|v|^2, v.z)
// run over a 16-lane butterfly, which gives the identical tree as the 32-lane one whose upper half adds zeros.
template <int CAP>
struct PairWS {  // per HALF-warp, in doubles
    static constexpr int LP = 0;                        // packed lower factor
    static constexpr int XS = CAP * (CAP + 1) / 2;      // x by passive position
    static constexpr int PI = XS + CAP;                 // passive list (ints)
    static constexpr int CS = PI + CAP / 2;             // c (16 NPH doubles follow)
};

__device__ __forceinline__ double hshfl(double v, int src)  // src: lane within the half
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(FULL, lo, src, 16);
    hi = __shfl_sync(FULL, hi, src, 16);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double hsum(double v)  // sum over the half (all its lanes receive it)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += shfl2_xor(v, o);
    return v;
}
// arg-max over the strictly positive values of a half (lanes with idx < 0 do not take part); ties -> lowest index
__device__ __forceinline__ void hargmax_pos(double &v, int &idx, unsigned hmask)
{
    const unsigned hi = idx >= 0 ? (unsigned)__double2hiint(v) : 0u, lo = idx >= 0 ? (unsigned)__double2loint(v) : 0u;
    const unsigned hm = __reduce_max_sync(hmask, hi);
    const unsigned lm = __reduce_max_sync(hmask, hi == hm ? lo : 0u);
    const bool win = idx >= 0 && hi == hm && lo == lm;
    const int widx = __reduce_min_sync(hmask, win ? idx : 0x7fffffff);
    v = __hiloint2double((int)hm, (int)lm);
    idx = (hm | lm) ? widx : -1;
}
// arg-min over a half (lanes with idx < 0 do not take part), ties -> lowest index
__device__ __forceinline__ void hargmin(double &v, int &idx, unsigned hmask)
{
    const unsigned long long k = idx >= 0 ? ~dkey(v) : 0ull;
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned hm = __reduce_max_sync(hmask, hi);
    const unsigned lm = __reduce_max_sync(hmask, hi == hm ? lo : 0u);
    const bool win = idx >= 0 && hi == hm && lo == lm;
    const int widx = __reduce_min_sync(hmask, win ? idx : 0x7fffffff);
    const unsigned long long km = ((unsigned long long)hm << 32) | lm;
    v = dkey_inv(~km);
    idx = km == 0ull ? -1 : widx;
}

// Deletion of passive position q (of pn) from the factor of each half that has `del` set (chol_delete for pairs).
__device__ __forceinline__ void chol_delete_pair(double *Lp, bool del, int q, int pn, double &rdl, double &zl, int h)
{
    const bool mine = del && (h >= q) && (h < pn - 1);
    const double *src = Lp + tri(h + 1, 0);  // old row h + 1 (only dereferenced when `mine`)
    double *dst = Lp + tri(h, 0);
    const int qd = del ? q : 0, len = del ? pn - 1 - q : 0;
    const int qmax = max(qd, __shfl_xor_sync(FULL, qd, 16)), lmax = max(len, __shfl_xor_sync(FULL, len, 16));
#pragma unroll 1
    for (int col = 0; col < qmax; ++col) {  // columns left of q: plain move (read everywhere before anyone overwrites)
        const bool on = mine && col < q;
        double lv = 0.0;
        if (on) lv = src[col];
        __syncwarp();
        if (on) dst[col] = lv;
    }
    double carry = 0.0;
    if (mine) carry = src[q];
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < lmax; ++i) {
        const int r = q + i;
        const bool step = i < len;
        const bool act = mine && step && h >= r;
        double u2 = 0.0;
        if (act) u2 = src[r + 1];
        const int rs = step ? r : 0;
        const double a = hshfl(carry, rs), b = hshfl(u2, rs);
        const double ir = rsqrt(fma(a, a, b * b));  // = 1 / (new diagonal element)
        const double cs = a * ir, sn = b * ir;
        __syncwarp();  // column r of row h + 1 was read one step ago; its owner may overwrite it now
        if (act) {
            dst[r] = fma(cs, carry, sn * u2);
            carry = fma(cs, u2, -sn * carry);
            if (h == r) rdl = ir;
        }
        const double zr = hshfl(zl, rs), zr1 = hshfl(zl, step ? r + 1 : 0);
        if (del && step) {
            if (h == r) zl = fma(cs, zr, sn * zr1);
            else if (h == r + 1) zl = fma(cs, zr1, -sn * zr);
        }
    }
    if (del && h >= pn - 1) zl = 0.0;
    __syncwarp();
}

// NNLS of two voxels (one per half-warp) on the full dictionary.  `have`: this half has a voxel.  hw: the half's workspace
// (PairWS), cs = hw + CS holds c.  Out (per half): zz, x_last / x_prev (atoms n - 1, n - 2), returns the overflow flag.
template <int NPH, int CAP>
__device__ __forceinline__ int nnls_pair(const double *__restrict__ T, const int ldT, const int n, const int mcap, const int itmax,
                                         double *hw, const bool have, const int lane, const int cap, const bool use_as, const ASpace as,
                                         double &zz, double &x_last, double &x_prev)
{
    const int h = lane & 15;
    const unsigned hmask = 0xffffu << (lane & 16);
    double *Lp = hw + PairWS<CAP>::LP, *xs = hw + PairWS<CAP>::XS, *cs = hw + PairWS<CAP>::CS;
    int *P = (int *)(hw + PairWS<CAP>::PI);
    int np = 0, iter = 0, overflow = 0;
    unsigned inP = 0u, avail = 0u;
    double xp = 0.0, zl = 0.0, rdl = 0.0;
    int myP = 0;
    bool run = have;
#pragma unroll
    for (int s = 0; s < NPH; ++s) avail |= (h + 16 * s < n ? 1u : 0u) << s;
    const double *Tl = T + h;
    const double *myrow = Lp + tri(h, 0);
    for (;;) {
        if (run && np >= mcap) run = false;
        if (run && np >= cap) { overflow = 1; run = false; }
        if (!__any_sync(FULL, run)) break;
        // ---- dual w = c - T[:,P] x_P in passive order (a finished / shorter half multiplies row 0 by zero)
        double wl[NPH];
#pragma unroll
        for (int s = 0; s < NPH; ++s) wl[s] = cs[h + 16 * s];
        {
            const int npr = run ? np : 0;
            const int npmax = max(npr, __shfl_xor_sync(FULL, npr, 16));
#pragma unroll 1
            for (int k = 0; k < npmax; k += 2) {
                const bool on0 = k < npr, on1 = k + 1 < npr;
                const double *r0 = Tl + (unsigned)((on0 ? P[k] : 0) * ldT), *r1 = Tl + (unsigned)((on1 ? P[k + 1] : 0) * ldT);
                const double x0 = on0 ? xs[k] : 0.0, x1 = on1 ? xs[k + 1] : 0.0;
                constexpr int HC = (NPH + 1) / 2;
#pragma unroll
                for (int c0 = 0; c0 < NPH; c0 += HC) {
                    double g0[HC], g1[HC];
#pragma unroll
                    for (int s = 0; s < HC; ++s) if (c0 + s < NPH) g0[s] = r0[16 * (c0 + s)];
#pragma unroll
                    for (int s = 0; s < HC; ++s) if (c0 + s < NPH) g1[s] = r1[16 * (c0 + s)];
#pragma unroll
                    for (int s = 0; s < HC; ++s) if (c0 + s < NPH) wl[c0 + s] = fma(-g0[s], x0, wl[c0 + s]);
#pragma unroll
                    for (int s = 0; s < HC; ++s) if (c0 + s < NPH) wl[c0 + s] = fma(-g1[s], x1, wl[c0 + s]);
                }
            }
        }
        // ---- candidate selection
        unsigned valid = avail & ~inP;
        bool pend = run, acc = false;
        int j = -1;
        double v = 0.0, d2 = 0.0, znum = 0.0;
        for (;;) {
            double bv = 0.0;
            int bj = -1;
            if (pend) {
#pragma unroll
                for (int s = 0; s < NPH; ++s)
                    if (((valid >> s) & 1u) && wl[s] > bv) { bv = wl[s]; bj = h + 16 * s; }
            }
            hargmax_pos(bv, bj, hmask);
            if (pend) {
                j = bj;
                if (j < 0) { pend = false; run = false; }  // no positive dual left: this voxel is done
            }
            if (!__any_sync(FULL, pend)) break;
            const int jj = pend ? j : 0;
            const double *Tj = T + (unsigned)(jj * ldT);
            double t = (pend && h < np) ? Tj[myP] : 0.0;  // = T[P[h]][j] (the table is exactly symmetric)
            const double hjj = Tj[jj], cj = cs[jj];
            const int npp = pend ? np : 0;
            const int npmax = max(npp, __shfl_xor_sync(FULL, npp, 16));
#pragma unroll 1
            for (int k = 0; k < npmax; ++k) {  // v = L^-1 t
                const double vk = hshfl(t * rdl, k);
                if (h > k && h < npp) t = fma(-myrow[k], vk, t);
            }
            v = (h < npp) ? t * rdl : 0.0;
            const double vv = hsum(v * v), vz = hsum(v * zl);
            d2 = hjj - vv;
            znum = cj - vz;
            const bool near = use_as && pend && np > 0 && d2 < 1e-10 * hjj;
            if (__any_sync(FULL, near)) {
                // near-dependent candidate: both quantities re-evaluated in A-space (see warp_nnls for the why)
                double beta = v;
                {
                    const int npn = near ? np : 0;
                    const int nmax = max(npn, __shfl_xor_sync(FULL, npn, 16));
                    const double *col = Lp + tri(nmax - 1, 0) + h;
#pragma unroll 1
                    for (int k = nmax - 1; k >= 0; --k) {
                        const double sk = hshfl(beta * rdl, k);
                        if (h < k && k < npn) beta = fma(-*col, sk, beta);
                        col -= k;
                    }
                    beta = (h < npn) ? beta * rdl : 0.0;
                    const float *Sj = as.S + jj;
                    double a2 = 0.0, ay = 0.0;
#pragma unroll 1
                    for (int i0 = 0; i0 < as.m; i0 += 16) {
                        const int i = i0 + h;
                        const bool on = i < as.m;
                        const float *Si = as.S + (size_t)(on ? i : 0) * as.n_pad;
                        double r = (double)Sj[(size_t)(on ? i : 0) * as.n_pad];
#pragma unroll 1
                        for (int a = 0; a < nmax; ++a) {
                            const float s0 = Si[a < npn ? P[a] : 0];
                            const double b0 = hshfl(beta, a);  // 0 beyond npn
                            r = fma(-(double)s0, b0, r);
                        }
                        if (on && near) {
                            const double yi = as.y_f64 ? ((const double *)as.y)[as.vox * as.m + i] : (double)((const float *)as.y)[as.vox * as.m + i];
                            a2 = fma(r, r, a2);
                            ay = fma(r, yi, ay);
                        }
                    }
                    a2 = hsum(a2);
                    ay = hsum(ay);
                    if (near) {
                        d2 = a2;
                        znum = ay;
                        if (d2 < 1e-24 * hjj) d2 = 0.0;
                    }
                }
            }
            if (pend) {
                if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) { acc = true; pend = false; }
                else if ((j & 15) == h) valid &= ~(1u << (j >> 4));
            }
            if (!__any_sync(FULL, pend)) break;
        }
        // ---- accepted candidates join the passive set: one more row of the factor
        if (acc) {
            const double ird = rsqrt(d2), dd = d2 * ird;
            double *rownew = Lp + tri(np, 0);
            if (h < np) rownew[h] = v;
            if (h == np) {
                rownew[h] = dd;
                rdl = ird;
                zl = znum * ird;
                xp = 0.0;
                myP = j;
                P[h] = j;
            }
            if ((j & 15) == h) inP |= 1u << (j >> 4);
            ++np;
        }
        __syncwarp();
        // ---- secondary loop
        bool need = acc;
        for (;;) {
            if (need && ++iter > itmax) { need = false; run = false; }
            if (!__any_sync(FULL, need)) break;
            double s;
            {   // s = L^-T z
                const int npn = need ? np : 0;
                const int nmax = max(npn, __shfl_xor_sync(FULL, npn, 16));
                s = (h < npn) ? zl : 0.0;
                const double *col = Lp + tri(nmax - 1, 0) + h;
#pragma unroll 1
                for (int k = nmax - 1; k >= 0; --k) {
                    const double sk = hshfl(s * rdl, k);
                    if (h < k && k < npn) s = fma(-*col, sk, s);
                    col -= k;
                }
                s = (h < npn) ? s * rdl : 0.0;
            }
            const bool neg = need && (h < np) && (s <= 0.0);
            const bool hneg = (__ballot_sync(FULL, neg) & hmask) != 0u;
            double tmin = INFINITY;
            int cand = -1;
            if (neg) {
                const double tt = -xp / (s - xp);
                if (tt < 2.0) { tmin = tt; cand = h; }
            }
            if (__any_sync(FULL, hneg)) hargmin(tmin, cand, hmask);
            if (need && (!hneg || cand < 0)) {  // feasible (or no admissible step): the solve stands
                if (h < np) { xp = s; xs[h] = s; }
                need = false;
            }
            if (!__any_sync(FULL, need)) break;
            // step back to the boundary, drop the coefficients that reached zero
            if (need) {
                if (h < np) xp = fma(tmin, s - xp, xp);
                if (h == cand) xp = 0.0;
            }
            const bool keep = need && (h < np) && (xp > 0.0);
            const unsigned kmask = (__ballot_sync(FULL, keep) >> (lane & 16)) & 0xffffu;
            const int np_old = np;
            unsigned rmask = need ? (~kmask & ((1u << np_old) - 1u)) : 0u;  // removed positions
            {
                const unsigned rany = rmask | __shfl_xor_sync(FULL, rmask, 16);
                for (unsigned r2 = rany; r2; r2 &= r2 - 1) {  // their atoms leave the passive bit set
                    const int pos = __ffs(r2) - 1;
                    const int a = __shfl_sync(FULL, myP, pos, 16);
                    if (((rmask >> pos) & 1u) && (a & 15) == h) inP &= ~(1u << (a >> 4));
                }
            }
            if (need) np = __popc(kmask);
            if (keep) {
                const int rank = __popc(kmask & ((1u << h) - 1u));
                P[rank] = myP;
                xs[rank] = xp;
            }
            __syncwarp();
            if (need) {
                if (h < np) { myP = P[h]; xp = xs[h]; }
                else xp = 0.0;
            }
            // Cholesky downdate, one removed position per half at a time, highest first
            {
                int pn = np_old;
                for (;;) {
                    const bool del = rmask != 0u;
                    if (!__any_sync(FULL, del)) break;
                    const int q = del ? 31 - __clz(rmask) : 0;
                    rmask &= ~(1u << q);
                    chol_delete_pair(Lp, del && np > 0, q, pn, rdl, zl, h);
                    --pn;
                }
            }
            if (need && np == 0) need = false;
        }
    }
    {
        const unsigned m1 = (__ballot_sync(FULL, h < np && myP == n - 1) >> (lane & 16)) & 0xffffu;
        const unsigned m2 = (__ballot_sync(FULL, h < np && myP == n - 2) >> (lane & 16)) & 0xffffu;
        const double x1 = hshfl(xp, m1 ? __ffs(m1) - 1 : 0), x2 = hshfl(xp, m2 ? __ffs(m2) - 1 : 0);
        x_last = m1 ? x1 : 0.0;
        x_prev = m2 ? x2 : 0.0;
    }
    zz = hsum(h < np ? zl * zl : 0.0);
    return overflow;
}

// NODDI stage 1, two voxels per warp (half-warp per voxel); same batch queue and outputs as k_noddi_stage<1>.
template <int NPH, int MAXT, int CAP>
__global__ void __launch_bounds__(MAXT, 1) k_noddi_stage1_pair(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int HSIZE = PairWS<CAP>::CS + 16 * NPH;          // per half
    double *wsb = (double *)(smem + p.ws_smem_off) + warp * (2 * HSIZE + BV);
    double *hw = wsb + (lane >> 4) * HSIZE, *bx = wsb + 2 * HSIZE;
    constexpr int NPL = (NPH + 1) / 2, NT = 4 * NPL, TP = (MAXT > 512 && NT % 2 == 0) ? NT / 2 : NT;
    const int n = p.n, NA = p.NA, h = lane & 15;
    const int cap = min(min(p.cap_stage[0], CAP), c_lc_cap);
    double *scr = p.scratch + ((size_t)blockIdx.x * 32 + warp) * (size_t)BV * NA;
    const int g = lane >> 2;
    int *counter = p.tile_counter;
    const int n_tiles = *p.n_tiles_ptr;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1);
        b = __reduce_add_sync(FULL, b);
        if (b >= n_tiles) break;
        const int4 tile = p.tiles[b];
        const int nb = tile.z;  // <= BV
        const float *S = (const float *)p.slab + (size_t)tile.x * p.slab_stride;
        const bool vvalid = g < nb;
        const long long mypos = tile.y + (vvalid ? g : 0);
        const long long myvox = (long long)p.order[mypos];
        const double *T1 = p.T1 + (size_t)tile.x * p.T1_stride;
        double *c1b = p.c1_all ? p.c1_all + (size_t)tile.y * NA : scr;
        gemm_c1<NT, TP, float>(S, p.n_pad, p.m, p.y, p.y_f64, myvox, vvalid, c1b, NA, lane, bx);
#pragma unroll 1
        for (int v0 = 0; v0 < nb; v0 += 2) {
            const int v = v0 + (lane >> 4);
            const bool have = v < nb;
            const long long pos = tile.y + (have ? v : v0);
            const long long vox = (long long)p.order[pos];
            const ASpace asp{S, p.n_pad, p.m, p.y, p.y_f64, vox};
            double *cs = hw + PairWS<CAP>::CS;
#pragma unroll
            for (int s = 0; s < NPH; ++s) cs[h + 16 * s] = (h + 16 * s < NA) ? c1b[(size_t)(have ? v : v0) * NA + h + 16 * s] : 0.0;
            __syncwarp();
            double zz, x1, x2;
            const int ov = nnls_pair<NPH, CAP>(T1, p.ldT1, n, p.m, 3 * n, hw, have, lane, cap, p.aspace != 0, asp, zz, x1, x2);
            if (h == 0 && have) {
                p.xiso[2 * pos] = x1;
                p.xiso[2 * pos + 1] = p.exvivo ? x2 : 0.0;
                const double yy = bx[v];
                if (yy > 0.0 && yy - zz < p.exact_tol * yy) {  // exact-fit voxel: queued for the A-space QR path (see k_noddi_stage)
                    const unsigned long long idx = atomicAdd((unsigned long long *)&p.status[4], 1ull);
                    if ((long long)idx < p.exact_cap) p.exact_list[idx] = (int)vox;
                }
            }
            if (ov && have) {  // queue_slow for one half
                if (h == 0) {
                    const unsigned long long idx = atomicAdd((unsigned long long *)&p.status[2], 1ull);
                    if ((long long)idx < p.ovf_cap) p.ovf_list[idx] = (int)vox;
                }
            }
            __syncwarp();
        }
    }
}

// ================================================================================================
// NODDI stage 3 (debias on the support, amico/models.pyx:929-942 + maps :945-979), ONE VOXEL PER THREAD.
// The system is tiny -- ~11 support atoms, never more than 6 passive ones on the default grids -- and a warp solving one
// such voxel spends ~4.4 k instructions mostly on cross-lane glue.  Here every thread runs the complete Lawson-Hanson
// iteration of warp_nnls<1, MAPPED> for its own voxel: same pivoting rules and the same floating-point operations in the same
// order (column-oriented substitutions, the butterfly trees of the two short sums written out for <= 8 terms, the same Givens
// downdates), so the coefficients are bit-identical.  Voxels are taken in LUT-direction order: the threads of a warp read the
// same Gram table / dictionary slab (L1 broadcast).  The one long step, the A-space re-evaluation of a near-dependent candidate
// (m x |P| products), is served COOPERATIVELY: the warp collects the requesting threads by ballot and evaluates one request
// at a time with all 32 lanes, rows strided over the lanes exactly as warp_nnls does it.
// A voxel whose passive set would outgrow CAPT (or whose support exceeds 32 atoms) is appended to `redo` as a one-voxel
// tile and re-fitted by k_noddi_stage<3>.  Per-thread state lives in shared memory as [index][thread].
constexpr int TPV_THREADS = 128;
template <int CAPT>
__host__ __device__ constexpr int tpv_smem_bytes()
{
    return TPV_THREADS * ((CAPT * (CAPT + 1) / 2 + 4 * CAPT) * 8 + CAPT * 4 + 32);
}

template <int NPL, int CAPT>
__global__ void __launch_bounds__(TPV_THREADS) k_noddi_stage3_tpv(const FitParams p, int4 *redo, int *redo_count)
{
    constexpr int NTH = TPV_THREADS, TRI_T = CAPT * (CAPT + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*Ls)[NTH] = (double (*)[NTH])smem_raw;          // packed lower factor
    double (*rds)[NTH] = Ls + TRI_T;                         // 1 / diagonal
    double (*zs)[NTH] = rds + CAPT;                          // z = L^-1 c_P
    double (*xs)[NTH] = zs + CAPT;                           // coefficients by passive position
    double (*bs)[NTH] = xs + CAPT;                           // scratch: v / s / beta
    int (*Ps)[NTH] = (int (*)[NTH])(bs + CAPT);              // passive list (compact indices)
    unsigned char (*sa)[NTH] = (unsigned char (*)[NTH])(Ps + CAPT);  // compact index -> atom
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = p.n, n_wm = p.n_wm, ldT = p.ldT1, NA = p.NA, m = p.m;
    const bool use_as = p.aspace != 0;
    for (long long base = (long long)blockIdx.x * NTH; base < p.n_vox; base += (long long)gridDim.x * NTH) {
        const long long pos = base + tid;
        const bool active = pos < p.n_vox;
        const long long vox = active ? (long long)p.order[pos] : 0;
        const int dir = active ? p.lut[vox] : 0;
        const double *T = p.T1 + (size_t)dir * p.T1_stride;
        const double *cg = p.c1_all + (size_t)(active ? pos : 0) * NA;
        int ns = 0;
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            unsigned w = active ? p.supmask[(size_t)pos * NPL + s] : 0u;
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                if (ns < 32) sa[ns][tid] = (unsigned char)(32 * s + b);
                ++ns;
            }
        }
        bool redo_me = active && ns > 32;
        bool run = active && !redo_me && ns > 0;
        const int itmax = 3 * ns;
        int np = 0, iter = 0;
        unsigned inP = 0u;  // bit a: compact index a is passive
        const unsigned allmask = ns >= 32 ? 0xffffffffu : ((1u << ns) - 1u);
        while (__any_sync(FULL, run)) {
            // ---- candidate selection (per thread), with the warp-cooperative A-space service inside
            bool pend = run && np < m;
            if (pend && np >= CAPT) { pend = false; run = false; redo_me = true; }
            if (run && !pend) run = false;  // np reached m: done
            unsigned valid = allmask & ~inP;
            bool acc = false;
            int j = -1;
            double d2 = 0.0, znum = 0.0, vv = 0.0, hjj = 0.0;
            while (__any_sync(FULL, pend)) {
                bool near = false;
                if (pend) {
                    // dual w = c - T[:,P] x_P over the still valid atoms, largest positive one (ties: lowest index)
                    double bv = 0.0;
                    j = -1;
                    for (unsigned mm = valid; mm; mm &= mm - 1) {
                        const int a = __ffs(mm) - 1, atom = sa[a][tid];
                        double w = cg[atom];
                        for (int k = 0; k < np; ++k) w = fma(-T[(unsigned)(sa[Ps[k][tid]][tid] * ldT + atom)], xs[k][tid], w);
                        if (w > bv) { bv = w; j = a; }
                    }
                    if (j < 0) { pend = false; run = false; }
                }
                if (pend) {
                    const int atomj = sa[j][tid];
                    const double *Tj = T + (unsigned)(atomj * ldT);
                    hjj = Tj[atomj];
                    // v = L^-1 t, t_k = T[P_k][j]: column-oriented forward substitution (the order warp_nnls uses)
                    for (int k = 0; k < np; ++k) bs[k][tid] = Tj[sa[Ps[k][tid]][tid]];
                    for (int k = 0; k < np; ++k) {
                        const double vk = bs[k][tid] * rds[k][tid];
                        for (int a = k + 1; a < np; ++a) bs[a][tid] = fma(-Ls[tri(a, k)][tid], vk, bs[a][tid]);
                    }
                    // |v|^2 and v.z over the 32-lane butterfly tree (offsets 16, 8, 4, 2, 1) restricted to its <= 8 non-zero leaves
                    double sq[8], sz[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool on = k < np && k < CAPT;
                        const double vk = on ? bs[k < CAPT ? k : 0][tid] * rds[k < CAPT ? k : 0][tid] : 0.0;
                        if (on) bs[k < CAPT ? k : 0][tid] = vk;  // bs now holds v
                        sq[k] = vk * vk;
                        sz[k] = vk * (on ? zs[k < CAPT ? k : 0][tid] : 0.0);
                    }
                    vv = ((sq[0] + sq[4]) + (sq[2] + sq[6])) + ((sq[1] + sq[5]) + (sq[3] + sq[7]));
                    const double vz = ((sz[0] + sz[4]) + (sz[2] + sz[6])) + ((sz[1] + sz[5]) + (sz[3] + sz[7]));
                    d2 = hjj - vv;
                    znum = cg[atomj] - vz;
                    near = use_as && np > 0 && d2 < 1e-10 * hjj;
                    if (near) {  // beta = L^-T v for the service below (column-oriented back substitution), into bs
                        for (int k = np - 1; k >= 0; --k) {
                            const double sk = bs[k][tid] * rds[k][tid];
                            bs[k][tid] = sk;
                            for (int a = 0; a < k; ++a) bs[a][tid] = fma(-Ls[tri(k, a)][tid], sk, bs[a][tid]);
                        }
                    }
                }
                // ---- A-space service: r = a_j - A_P beta, d2 = |r|^2, numerator r.y, one requesting thread at a time
                for (unsigned req = __ballot_sync(FULL, near); req; req &= req - 1) {
                    const int src = __ffs(req) - 1, st = (tid & ~31) + src;
                    const int r_np = __shfl_sync(FULL, np, src), r_j = __shfl_sync(FULL, j, src), r_dir = __shfl_sync(FULL, dir, src);
                    const long long r_vox = __shfl_sync(FULL, vox, src);
                    const float *S = (const float *)p.slab + (size_t)r_dir * p.slab_stride;
                    const float *Sj = S + sa[r_j][st];
                    double a2 = 0.0, ay = 0.0;
#pragma unroll 1
                    for (int i0 = 0; i0 < m; i0 += 32) {
                        const int i = i0 + lane;
                        const bool on = i < m;
                        const float *Si = S + (size_t)(on ? i : 0) * p.n_pad;
                        double r = (double)Sj[(size_t)(on ? i : 0) * p.n_pad];
#pragma unroll 1
                        for (int a = 0; a < r_np; a += 2) {
                            const int a1 = min(a + 1, r_np - 1);
                            const float s0 = Si[sa[Ps[a][st]][st]], s1 = Si[sa[Ps[a1][st]][st]];
                            const double b0 = bs[a][st], b1 = (a + 1 < r_np) ? bs[a1][st] : 0.0;
                            r = fma(-(double)s0, b0, r);
                            r = fma(-(double)s1, b1, r);
                        }
                        if (on) {
                            const double yi = p.y_f64 ? ((const double *)p.y)[r_vox * m + i] : (double)((const float *)p.y)[r_vox * m + i];
                            a2 = fma(r, r, a2);
                            ay = fma(r, yi, ay);
                        }
                    }
                    a2 = warp_sum(a2);
                    ay = warp_sum(ay);
                    if (lane == src) {
                        d2 = a2;
                        znum = ay;
                        if (d2 < 1e-24 * hjj) d2 = 0.0;
                    }
                }
                if (pend) {
                    if (near) {  // bs held beta: restore v = L^-1 t for the new factor row (same operations as above)
                        const double *Tj = T + (unsigned)(sa[j][tid] * ldT);
                        for (int k = 0; k < np; ++k) bs[k][tid] = Tj[sa[Ps[k][tid]][tid]];
                        for (int k = 0; k < np; ++k) {
                            const double vk = bs[k][tid] * rds[k][tid];
                            for (int a = k + 1; a < np; ++a) bs[a][tid] = fma(-Ls[tri(a, k)][tid], vk, bs[a][tid]);
                            bs[k][tid] = vk;
                        }
                    }
                    if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) { acc = true; pend = false; }
                    else valid &= ~(1u << j);
                }
            }
            if (!acc) continue;
            // ---- j joins the passive set (from here on purely per thread)
            {
                const double ird = rsqrt(d2), dd = d2 * ird;
                for (int k = 0; k < np; ++k) Ls[tri(np, k)][tid] = bs[k][tid];
                Ls[tri(np, np)][tid] = dd;
                rds[np][tid] = ird;
                zs[np][tid] = znum * ird;
                xs[np][tid] = 0.0;
                Ps[np][tid] = j;
                inP |= 1u << j;
                ++np;
            }
            for (;;) {  // secondary loop
                if (++iter > itmax) { run = false; break; }
                // s = L^-T z into bs (column-oriented back substitution)
                for (int k = 0; k < np; ++k) bs[k][tid] = zs[k][tid];
                for (int k = np - 1; k >= 0; --k) {
                    const double sk = bs[k][tid] * rds[k][tid];
                    bs[k][tid] = sk;
                    for (int a = 0; a < k; ++a) bs[a][tid] = fma(-Ls[tri(k, a)][tid], sk, bs[a][tid]);
                }
                double tmin = INFINITY;
                int cand = -1;
                bool anyneg = false;
                for (int k = 0; k < np; ++k) {
                    const double sk = bs[k][tid];
                    if (sk <= 0.0) {
                        anyneg = true;
                        const double xk = xs[k][tid], tt = -xk / (sk - xk);
                        if (tt < 2.0 && tt < tmin) { tmin = tt; cand = k; }
                    }
                }
                if (!anyneg || cand < 0) {
                    for (int k = 0; k < np; ++k) xs[k][tid] = bs[k][tid];
                    break;
                }
                unsigned rmask = 0u;
                const int np_old = np;
                int q = 0;
                for (int k = 0; k < np_old; ++k) {
                    double xk = xs[k][tid];
                    xk = fma(tmin, bs[k][tid] - xk, xk);
                    if (k == cand) xk = 0.0;
                    if (xk > 0.0) {
                        xs[q][tid] = xk;
                        Ps[q][tid] = Ps[k][tid];
                        ++q;
                    } else {
                        rmask |= 1u << k;
                        inP &= ~(1u << Ps[k][tid]);
                    }
                }
                np = q;
                if (np == 0) break;
                for (int pn = np_old; rmask; --pn) {  // Cholesky downdate, highest removed position first
                    const int qd = 31 - __clz(rmask);
                    rmask &= ~(1u << qd);
                    // rows qd+1.. move up one slot, Givens rotations restore the triangle, z is rotated along (chol_delete)
                    double carry[CAPT];
#pragma unroll
                    for (int i = 0; i < CAPT; ++i) carry[i] = 0.0;
                    for (int i = qd; i < pn - 1; ++i)
                        for (int col = 0; col < qd; ++col) Ls[tri(i, col)][tid] = Ls[tri(i + 1, col)][tid];
#pragma unroll
                    for (int i = 0; i < CAPT; ++i)
                        if (i >= qd && i < pn - 1) carry[i] = Ls[tri(i + 1, qd)][tid];
                    for (int r = qd; r < pn - 1; ++r) {
                        double a = 0.0;
#pragma unroll
                        for (int i = 0; i < CAPT; ++i)
                            if (i == r) a = carry[i];
                        const double b = Ls[tri(r + 1, r + 1)][tid];
                        const double ir = rsqrt(fma(a, a, b * b));
                        const double cs = a * ir, sn = b * ir;
#pragma unroll
                        for (int i = 0; i < CAPT; ++i) {
                            if (i >= r && i < pn - 1) {
                                const double u2 = Ls[tri(i + 1, r + 1)][tid];
                                Ls[tri(i, r)][tid] = fma(cs, carry[i], sn * u2);
                                carry[i] = fma(cs, u2, -sn * carry[i]);
                            }
                        }
                        rds[r][tid] = ir;
                        const double zr = zs[r][tid], zr1 = zs[r + 1][tid];
                        zs[r][tid] = fma(cs, zr, sn * zr1);
                        zs[r + 1][tid] = fma(cs, zr1, -sn * zr);
                    }
                    zs[pn - 1][tid] = 0.0;
                }
            }
        }
        // ---- maps (noddi_maps: sums over the positive coefficients in ascending atom order)
        if (active && !redo_me) {
            auto xof = [&](int a) {  // coefficient of compact index a (passive atoms only)
                double xv = 0.0;
                for (int k = 0; k < np; ++k)
                    if (Ps[k][tid] == a) xv = xs[k][tid];
                return xv;
            };
            double s_all = 0.0;
            for (unsigned mm = inP; mm; mm &= mm - 1) {
                const double xj = xof(__ffs(mm) - 1);
                if (xj > 0.0) s_all += xj;
            }
            s_all += 1e-16;
            double s_wm = 0.0;
            for (unsigned mm = inP; mm; mm &= mm - 1) {
                const int a = __ffs(mm) - 1;
                const double xj = xof(a);
                if (xj > 0.0 && sa[a][tid] < n_wm) s_wm += xj / s_all;
            }
            s_wm += 1e-16;
            double f1 = 0.0, f2 = 0.0, k1 = 0.0;
            for (unsigned mm = inP; mm; mm &= mm - 1) {
                const int a = __ffs(mm) - 1, atom = sa[a][tid];
                const double xj = xof(a);
                if (xj > 0.0 && atom < n_wm) {
                    const float ic = p.icvf[atom];
                    f1 += (double)ic * xj / s_all / s_wm;
                    f2 += (double)((float)(1.0 - (double)ic)) * xj / s_all / s_wm;
                    k1 += (double)p.kappa[atom] * xj / s_all / s_wm;
                }
            }
            const double ndi = f1 / (f1 + f2 + 1e-16);
            const double odi = 2.0 / 3.14159265358979323846 * atan2(1.0, k1);
            const double x_iso = ((inP >> (ns - 1)) & 1u) ? xof(ns - 1) : 0.0;  // atom n - 1 is always the last support member
            const double fwf = x_iso / s_all;
            double *e = p.est + vox * p.n_maps;
            e[0] = ndi; e[1] = odi; e[2] = fwf;
            if (p.exvivo) e[3] = (ns >= 2 && ((inP >> (ns - 2)) & 1u) ? xof(ns - 2) : 0.0) / s_all;
            if (p.flags & FLAG_EXTRA) {
                const double tf = 1.0 - fwf;
                p.extra[2 * vox] = ndi * tf;
                p.extra[2 * vox + 1] = odi * tf;
            }
            if (p.support_out) p.support_out[vox] = ns;
        }
        if (redo_me) {
            const int slot = atomicAdd(redo_count, 1);
            redo[slot] = make_int4(dir, (int)pos, 1, 0);
        }
        __syncwarp();
    }
}

}  // namespace amx
