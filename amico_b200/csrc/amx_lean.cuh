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

// 64-bit shuffles: amx_warp.cuh (two 32-bit shuffles, no volatile asm pack)
__device__ __forceinline__ double shfl2(double v, int src) { return shfl(v, src); }
__device__ __forceinline__ double shfl2_xor(double v, int m) { return shfl_xor(v, m); }

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
    asm volatile("" : "+r"(avail));  // opaque: otherwise re-derived from lane and n in every outer iteration (3 % of the instructions)
    const double *Tl = T + lane;
    asm volatile("" : "+l"(Tl));     // likewise the table pointer (re-built from the kernel parameters before every row otherwise)
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
            const double *Tj = Tl + (j * ldT - lane);  // = T + j ldT, from the pointer that is held in registers
            double t = (lane < np) ? Tj[myP] : 0.0;  // = T[P[lane]][j] (the table is exactly symmetric)
            const double hjj = Tj[j], cj = cs[j];
            {   // v = L^-1 t (forward substitution, lane a owns row a); two steps per trip: half the loop / convergence-check overhead
                const bool mine = lane < np;
                int k = 0;
#pragma unroll 1
                for (; k + 2 <= np; k += 2) {
                    const double v0 = shfl2(t * rdl, k);
                    if (mine && lane > k) t = fma(-myrow[k], v0, t);
                    const double v1 = shfl2(t * rdl, k + 1);
                    if (mine && lane > k + 1) t = fma(-myrow[k + 1], v1, t);
                }
                if (k < np) {
                    const double v0 = shfl2(t * rdl, k);
                    if (mine && lane > k) t = fma(-myrow[k], v0, t);
                }
            }
            v = (lane < np) ? t * rdl : 0.0;
            double vv = v * v, vz = v * zl;  // both 0 beyond np
            if (np > 0) warp_sum2_upto(vv, vz, np);
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
                int k = np - 1;
#pragma unroll 1
                for (; k >= 1; k -= 2) {  // two steps per trip (step 0 updates no lane)
                    const double s0 = shfl2(s * rdl, k);
                    if (lane < k) s = fma(-*col, s0, s);
                    col -= k;
                    const double s1 = shfl2(s * rdl, k - 1);
                    if (lane < k - 1) s = fma(-*col, s1, s);
                    col -= k - 1;
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
    int lane = threadIdx.x & 31;
    unsigned wso = p.ws_smem_off + (threadIdx.x >> 5) * ((LeanWS<CAP>::SIZE + 32 * NPL) * 8);
    asm volatile("" : "+r"(lane), "+r"(wso));  // opaque: kept in registers instead of being re-derived from %tid at every use
    const int warp = threadIdx.x >> 5;
    double *wsb = (double *)(smem + wso);
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

// ------------------------------------------------------------------------------------------------
// NODDI stage 2 (support selection, amico/models.pyx:918-936) on an inlined copy of warp_lars_fast: same path, same rules,
// same arithmetic (identical support words), with what the stage-1 rewrite taught -- no call ABI around the Gram-row loads
// (the __noinline__ solver re-materialised its global-memory descriptor before every load), 32-bit offsets, 64-bit
// shuffles without the register swaps, no coefficient vector.
// sum_{c < n} M(r, c) v[c] for the packed symmetric matrix M (see sym_row_dot), with running indices instead of tri() per term
__device__ __forceinline__ double sym_row_dot_lean(const double *Mi, const int r, const int n, const double *v)
{
    double acc = 0.0;
    const double *prow = Mi + tri(r, 0);  // (r, c), c < r
    const double *pcol = Mi + r;          // (c, r), c >= r: at tri(c, 0) + r
    int c = 0;
#pragma unroll 1
    for (; c + 2 <= n; c += 2) {  // two terms per trip; (c, r) sits c + 1 packed entries after (c - 1, r)
        acc = fma(*(c < r ? prow : pcol), v[c], acc);
        pcol += c + 1;
        acc = fma(*(c + 1 < r ? prow + 1 : pcol), v[c + 1], acc);
        pcol += c + 2;
        prow += 2;
    }
    if (c < n) acc = fma(*(c < r ? prow : pcol), v[c], acc);
    return acc;
}

template <int NPL>
struct Lars2WS {  // per-warp workspace of stage 2 (doubles), active-set capacity LC
    static constexpr int DTR = 0, MI = 32 * NPL, U = MI + LC * (LC + 1) / 2, GS = U + LC, IND = GS + LC, BX = IND + LC / 2, SIZE = BX + BV;
};

template <int NPL>
__device__ __forceinline__ int lars_lean(const double *__restrict__ T, const int ldT, const int K, const int Ltrue, const double lambda1, double *wsb,
                                         double normX, const int lane, int cap, unsigned (&sup)[NPL])
{
    // one base, compile-time offsets (a struct of pointers was re-materialised from the kernel parameters at every use at 64 registers)
    double *DtR = wsb + Lars2WS<NPL>::DTR, *Mi = wsb + Lars2WS<NPL>::MI, *u = wsb + Lars2WS<NPL>::U, *gs = wsb + Lars2WS<NPL>::GS;
    int *ind = (int *)(wsb + Lars2WS<NPL>::IND);
    // Result: sup[s] bit l set <=> atom l + 32 s ends with a positive coefficient (all lanes hold the words).
    cap = min(cap, c_lc_cap);
    int L = Ltrue < K ? Ltrue : K;
    int overflow = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) sup[s] = 0u;
    if (L <= 0) return 0;
    unsigned kin = 0u;  // bit s: atom lane + 32 s exists
#pragma unroll
    for (int s = 0; s < NPL; ++s) kin |= (lane + 32 * s < K ? 1u : 0u) << s;
    asm volatile("" : "+r"(kin));  // opaque: kept in a register instead of being re-derived in every step
    const double *Tq = T;
    asm volatile("" : "+l"(Tq));   // likewise the table pointer
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
                g = Tq[(unsigned)(cur * ldT + ind_l)];
                gs[lane] = g;
            }
            __syncwarp();
            if (i == 0) {
                if (lane == 0) Mi[0] = 1.0 / g;
                rs_l = (lane == 0) ? 1.0 / g : 0.0;
            } else {
                double ur = 0.0;
                if (lane < i) {
                    ur = sym_row_dot_lean(Mi, lane, i, gs);
                    u[lane] = ur;
                }
                double dot = lane < i ? ur * g : 0.0, usum = lane < i ? ur : 0.0;  // two interleaved butterfly sums
                warp_sum2_upto(dot, usum, i);
                const double schur = 1.0 / (shfl2(g, i) - dot);
                // row sums of the inverse after the Schur update: old rows += schur u_r (sum(u) - 1), new row = schur (1 - sum(u))
                if (lane < i) rs_l = fma(schur * ur, usum - 1.0, rs_l);
                if (lane == i) rs_l = schur * (1.0 - usum);
                __syncwarp();
                if (lane < i) {
                    const double su = schur * ur;
                    double *mp = Mi + tri(lane, lane);  // element (k, lane), k = lane..: the packed index grows by k + 1 per row
                    int k = lane;
#pragma unroll 1
                    for (; k + 2 <= i; k += 2) {  // two rows per trip
                        mp[0] = fma(su, u[k], mp[0]);
                        mp[k + 1] = fma(su, u[k + 1], mp[k + 1]);
                        mp += 2 * k + 3;
                    }
                    if (k < i) {
                        *mp = fma(su, u[k], *mp);
                        mp += k + 1;
                    }
                    *mp = -su;  // = Mi[tri(i, lane)]
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
            ul = sym_row_dot_lean(Mi, lane, i + 1, gs);
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
        const double cc = fabs(shfl2(dl, 0));
        // correlation slopes T[:, S] u; rows are L2-resident: fetch GD rows at a time
        double sl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) sl[s] = 0.0;
        {   // two rows in flight, accumulated in path order
            const double *Tl = Tq + lane;
            int j = 0;
#pragma unroll 1
            for (; j + 2 <= i + 1; j += 2) {
                const double *r0 = Tl + (unsigned)(ind[j] * ldT), *r1 = Tl + (unsigned)(ind[j + 1] * ldT);
                double g0[NPL], g1[NPL];
#pragma unroll
                for (int s = 0; s < NPL; ++s) g0[s] = r0[32 * s];  // columns >= K: finite values of the next row / the padding,
#pragma unroll
                for (int s = 0; s < NPL; ++s) g1[s] = r1[32 * s];  // only ever combined into slots that are masked by k < K
                const double u0 = u[j], u1 = u[j + 1];
#pragma unroll
                for (int s = 0; s < NPL; ++s) sl[s] = fma(g0[s], u0, sl[s]);
#pragma unroll
                for (int s = 0; s < NPL; ++s) sl[s] = fma(g1[s], u1, sl[s]);
            }
            if (j <= i) {
                const double *r0 = Tl + (unsigned)(ind[j] * ldT);
                const double u0 = u[j];
#pragma unroll
                for (int s = 0; s < NPL; ++s) sl[s] = fma(r0[32 * s], u0, sl[s]);
            }
        }
        // first inactive atom reaching the common correlation: entry of smallest magnitude, lowest index.  Each lane first
        // picks the best of its own atoms by cross-multiplication (|a/b| < |c/d| <=> |a| d < |c| b for b, d > 0), so only one
        // reciprocal per lane and step is needed.
        double bnum = INFINITY, bden = 1.0;  // best candidate of this lane: step = bnum / bden (bden > 0); inf / 1 while there is none
        int mk = lane < K ? lane : -1;       // its atom; lanes without a candidate still offer their lowest atom (step = inf)
        const unsigned open_ = kin & ~act;  // bit s: atom lane + 32 s exists and is not active
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            if (((open_ >> s) & 1u) && sl[s] < 1.0) {
                const int k = lane + 32 * s;
                const double num = cc - DtR[k], den = 1.0 - sl[s];
                if (fabs(num) * bden < fabs(bnum) * den) { bnum = num; bden = den; mk = k; }  // (the first candidate beats inf / 1)
            }
        }
        const double mine = bnum * __drcp_rn(bden);
        double bt = fabs(mine);
        int bk = mk;
        warp_argmin_nonneg(bt, bk);
        const double step0 = shfl2(mine, bk & 31);  // lane (bk & 31) owns atom bk and offered exactly it
        double step = step0;
        cur = bk;
        double coeff1 = lane <= i ? sg * ul : 0.0, coeff2 = lane <= i ? dl * ul : 0.0;  // two interleaved butterfly sums
        warp_sum2_upto(coeff1, coeff2, i + 1);
        const double step_max2 = cc - lambda1;
        step = fmin(fmin(step, step_max2), step_max);
        if (step == INFINITY) break;
        if (lane <= i) {
            coef_l = fma(step, ul, coef_l);
            if (coef_l < 0.0) coef_l = 0.0;
        }
#pragma unroll
        for (int s = 0; s < NPL; ++s)
            if ((kin >> s) & 1u) DtR[lane + 32 * s] = fma(-step, sl[s], DtR[lane + 32 * s]);
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
            double cn = shfl2(coef_l, min(lane + 1, 31));
            int in_ = __shfl_down_sync(FULL, ind_l, 1);
            const double rn = shfl2(rs_l, min(lane + 1, 31));
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
    return overflow;
}


template <int NPL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_noddi_stage2_lean(const FitParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    int lane = threadIdx.x & 31;
    unsigned wso = p.ws_smem_off + (threadIdx.x >> 5) * (Lars2WS<NPL>::SIZE * 8);
    asm volatile("" : "+r"(lane), "+r"(wso));  // opaque: kept in registers instead of being re-derived from %tid at every use
    const int warp = threadIdx.x >> 5;
    const int cap = p.cap_stage[1];
    double *wsb = (double *)(smem + wso);
    double *bx = wsb + Lars2WS<NPL>::BX;
    constexpr int NT = 4 * NPL, TP = (MAXT > 512 && NT % 2 == 0) ? NT / 2 : NT;
    const int m = p.m, n = p.n, n_pad = p.n_pad, n_wm = p.n_wm, NA = p.NA;
    double *scr = p.scratch + ((size_t)blockIdx.x * 32 + warp) * (size_t)BV * NA;
    const int g = lane >> 2;
    int *counter = p.tile_counter + 1;
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
        const double *T2 = p.T2 + (size_t)tile.x * p.T2_stride;
        const double xi = p.xiso[2 * mypos], xd = p.xiso[2 * mypos + 1];
        if (p.norms_const)
            gemm_c2<NT, TP, float, true>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo, p.norms, n_wm, scr,
                                         NA, bx, lane);
        else
            gemm_c2<NT, TP, float, false>(S, n_pad, n, n_wm, p.dc, p.dwi_rows, p.y, p.y_f64, m, myvox, vvalid, xi, xd, p.exvivo, p.norms, n_wm, scr,
                                          NA, bx, lane);
#pragma unroll 1
        for (int v = 0; v < nb; ++v) {
#pragma unroll
            for (int s = 0; s < NPL; ++s) wsb[Lars2WS<NPL>::DTR + lane + 32 * s] = scr[(size_t)v * NA + lane + 32 * s];
            __syncwarp();
            unsigned sup[NPL];
            const int ov = lars_lean<NPL>(T2, p.ldT2, n_wm, p.dc < n_wm ? p.dc : n_wm, p.lambda1, wsb, bx[v], lane, cap, sup);
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const int j = lane + 32 * s;
                const unsigned w = sup[s] | __ballot_sync(FULL, j >= n_wm && j < n);  // dot / iso columns always belong
                if (lane == s) p.supmask[(size_t)(tile.y + v) * NPL + s] = w;
            }
            if (ov) queue_slow(p, (long long)p.order[tile.y + v], lane);
            __syncwarp();
        }
    }
}

// ================================================================================================
// ONE VOXEL PER THREAD with warp-cooperative services (NODDI stages 1 and 3).
// A warp solving ONE NNLS voxel spends most of its ~4-9 k instructions on cross-lane glue around a passive set of <= 8 atoms
// (substitution steps through shuffles, butterflies, factor updates, control) with most lanes idle.  Here every THREAD owns a
// voxel and runs the small part of Lawson-Hanson for it -- candidate test, factor append, passive solve, ratio test,
// Givens downdates -- with its state in shared memory as [index][thread]; 32 voxels advance per instruction.  The WIDE parts
// are served by the whole warp, one requesting thread at a time (requests collected by ballot, the requester's state read from
// its shared-memory column):
//   * stage 1: the dual pass w = c - T[:,P] x over the full dictionary and its arg-max (Gram rows read coalesced, lane l owns
//     atoms l, l + 32, ... exactly as warp_nnls does),
//   * both stages: the A-space re-evaluation of a near-dependent candidate (m x |P| products, rows strided over the lanes).
// Same pivoting rules and the same floating-point operations in the same order as warp_nnls (column-oriented substitutions, the
// butterfly trees of the two short sums written out for <= 8 terms, the same Givens downdates): bit-identical coefficients.
// Voxels are taken in LUT-direction order, so a warp's threads mostly share the Gram table / dictionary slab.
// A voxel whose passive set would outgrow CAPT (or, stage 3, whose support exceeds 32 atoms) is appended to `redo` as a
// one-voxel tile and re-fitted by the warp-per-voxel kernel of the stage.
constexpr int TPV_THREADS = 128;

template <int CAPT>
struct TpvWS {  // per-thread state, [index][thread]
    static constexpr int NTH = TPV_THREADS, TRI_T = CAPT * (CAPT + 1) / 2;
    double (*Ls)[NTH];   // packed lower factor
    double (*rds)[NTH];  // 1 / diagonal
    double (*zs)[NTH];   // z = L^-1 c_P
    double (*xs)[NTH];   // coefficients by passive position
    double (*bs)[NTH];   // scratch: t / v / s / beta
    unsigned char (*Ps)[NTH];  // passive list (atoms / compact indices < 256)
    static constexpr int BYTES_PER_THREAD = (TRI_T + 4 * CAPT) * 8 + 8;  // CAPT <= 8 list bytes
    __device__ __forceinline__ unsigned char *carve(unsigned char *base)
    {
        Ls = (double (*)[NTH])base;
        rds = Ls + TRI_T;
        zs = rds + CAPT;
        xs = zs + CAPT;
        bs = xs + CAPT;
        Ps = (unsigned char (*)[NTH])(bs + CAPT);
        return (unsigned char *)(Ps + 8);
    }
};

// Candidate test of atom `atomj` against the passive set (per thread): v = L^-1 t with t_k = T[P_k][j] (column-oriented forward
// substitution, left in ws.bs), |v|^2 and v.z over the 32-lane butterfly tree restricted to its <= 8 non-zero leaves,
// d2 = H_jj - |v|^2, znum = c_j - v.z.  ATOM(k): dictionary atom of passive position k.
template <int CAPT, typename ATOM>
__device__ __forceinline__ void tpv_candidate(const TpvWS<CAPT> &ws, const int tid, const int np, const double *Tj, const double hjj, const double cj,
                                              ATOM atom_of, double &vv, double &d2, double &znum)
{
    for (int k = 0; k < np; ++k) ws.bs[k][tid] = Tj[atom_of(k)];
    for (int k = 0; k < np; ++k) {
        const double vk = ws.bs[k][tid] * ws.rds[k][tid];
        for (int a = k + 1; a < np; ++a) ws.bs[a][tid] = fma(-ws.Ls[tri(a, k)][tid], vk, ws.bs[a][tid]);
    }
    double sq[8], sz[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool on = k < np && k < CAPT;
        const int kk = k < CAPT ? k : 0;
        const double vk = on ? ws.bs[kk][tid] * ws.rds[kk][tid] : 0.0;
        if (on) ws.bs[kk][tid] = vk;  // bs now holds v
        sq[k] = vk * vk;
        sz[k] = vk * (on ? ws.zs[kk][tid] : 0.0);
    }
    vv = ((sq[0] + sq[4]) + (sq[2] + sq[6])) + ((sq[1] + sq[5]) + (sq[3] + sq[7]));
    const double vz = ((sz[0] + sz[4]) + (sz[2] + sz[6])) + ((sz[1] + sz[5]) + (sz[3] + sz[7]));
    d2 = hjj - vv;
    znum = cj - vz;
}

// beta = L^-T v in place (column-oriented back substitution on ws.bs)
template <int CAPT>
__device__ __forceinline__ void tpv_back_inplace(const TpvWS<CAPT> &ws, const int tid, const int np)
{
    for (int k = np - 1; k >= 0; --k) {
        const double sk = ws.bs[k][tid] * ws.rds[k][tid];
        ws.bs[k][tid] = sk;
        for (int a = 0; a < k; ++a) ws.bs[a][tid] = fma(-ws.Ls[tri(k, a)][tid], sk, ws.bs[a][tid]);
    }
}

// A-space service (warp-cooperative): for every thread with `near`, r = a_j - A_P beta (beta in the requester's ws.bs),
// d2 = |r|^2, znum = r.y with the rows strided over the lanes and summed exactly as warp_nnls does it.
template <int CAPT, typename ATOMT>
__device__ __forceinline__ void tpv_aspace_service(const TpvWS<CAPT> &ws, const FitParams &p, const bool near, const int tid, const int lane,
                                                   const int np, const int atomj, const int dir, const long long vox, const double hjj,
                                                   ATOMT atom_of_thread, double &d2, double &znum)
{
    const int m = p.m;
    __syncwarp();  // the requesters' beta / passive lists are read by the other lanes
    for (unsigned req = __ballot_sync(FULL, near); req; req &= req - 1) {
        const int src = __ffs(req) - 1, st = (tid & ~31) + src;
        const int r_np = __shfl_sync(FULL, np, src), r_atomj = __shfl_sync(FULL, atomj, src), r_dir = __shfl_sync(FULL, dir, src);
        const long long r_vox = __shfl_sync(FULL, vox, src);
        const float *S = (const float *)p.slab + (size_t)r_dir * p.slab_stride;
        const float *Sj = S + r_atomj;
        double a2 = 0.0, ay = 0.0;
        // four row chunks (128 rows) per trip, every chunk's dictionary loads in flight together; each row's residual and the two
        // per-lane sums are accumulated in the order warp_nnls uses (passive atoms in pairs, rows lane, lane + 32, ...)
        constexpr int CH = 4;
#pragma unroll 1
        for (int i0 = 0; i0 < m; i0 += 32 * CH) {
            const float *Si[CH];
            double r[CH], yv[CH];
            bool on[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int i = i0 + 32 * c + lane;
                on[c] = i < m;
                const size_t row = (size_t)(on[c] ? i : 0) * p.n_pad;
                Si[c] = S + row;
                r[c] = (double)Sj[row];
                yv[c] = !on[c] ? 0.0 : p.y_f64 ? ((const double *)p.y)[r_vox * m + i] : (double)((const float *)p.y)[r_vox * m + i];
            }
#pragma unroll 1
            for (int a = 0; a < r_np; a += 2) {
                const int a1 = min(a + 1, r_np - 1);
                const int at0 = atom_of_thread(a, st), at1 = atom_of_thread(a1, st);
                const double b0 = ws.bs[a][st], b1 = (a + 1 < r_np) ? ws.bs[a1][st] : 0.0;
                float s0[CH], s1[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) { s0[c] = Si[c][at0]; s1[c] = Si[c][at1]; }
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    r[c] = fma(-(double)s0[c], b0, r[c]);
                    r[c] = fma(-(double)s1[c], b1, r[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (on[c]) {
                    a2 = fma(r[c], r[c], a2);
                    ay = fma(r[c], yv[c], ay);
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
}

// v = L^-1 t again (ws.bs held beta during the service): same operations as tpv_candidate
template <int CAPT, typename ATOM>
__device__ __forceinline__ void tpv_restore_v(const TpvWS<CAPT> &ws, const int tid, const int np, const double *Tj, ATOM atom_of)
{
    for (int k = 0; k < np; ++k) ws.bs[k][tid] = Tj[atom_of(k)];
    for (int k = 0; k < np; ++k) {
        const double vk = ws.bs[k][tid] * ws.rds[k][tid];
        for (int a = k + 1; a < np; ++a) ws.bs[a][tid] = fma(-ws.Ls[tri(a, k)][tid], vk, ws.bs[a][tid]);
        ws.bs[k][tid] = vk;
    }
}

// Candidate j (list entry `pj`) joins the passive set, then the secondary loop: solve on the passive set, step back to the
// feasible boundary while a coefficient is not positive, Givens downdates for the atoms that leave.  REMOVED(entry): called
// for every passive-list entry that leaves.  Returns false when the iteration cap stopped the voxel.
template <int CAPT, typename REMOVED>
__device__ __forceinline__ bool tpv_accept(const TpvWS<CAPT> &ws, const int tid, int &np, int &iter, const int itmax, const int pj,
                                           const double d2, const double znum, REMOVED removed)
{
    {
        const double ird = rsqrt(d2), dd = d2 * ird;
        for (int k = 0; k < np; ++k) ws.Ls[tri(np, k)][tid] = ws.bs[k][tid];
        ws.Ls[tri(np, np)][tid] = dd;
        ws.rds[np][tid] = ird;
        ws.zs[np][tid] = znum * ird;
        ws.xs[np][tid] = 0.0;
        ws.Ps[np][tid] = (unsigned char)pj;
        ++np;
    }
    for (;;) {
        if (++iter > itmax) return false;
        for (int k = 0; k < np; ++k) ws.bs[k][tid] = ws.zs[k][tid];
        tpv_back_inplace<CAPT>(ws, tid, np);  // s = L^-T z
        double tmin = INFINITY;
        int cand = -1;
        bool anyneg = false;
        for (int k = 0; k < np; ++k) {
            const double sk = ws.bs[k][tid];
            if (sk <= 0.0) {
                anyneg = true;
                const double xk = ws.xs[k][tid], tt = -xk / (sk - xk);
                if (tt < 2.0 && tt < tmin) { tmin = tt; cand = k; }
            }
        }
        if (!anyneg || cand < 0) {
            for (int k = 0; k < np; ++k) ws.xs[k][tid] = ws.bs[k][tid];
            return true;
        }
        unsigned rmask = 0u;
        const int np_old = np;
        int q = 0;
        for (int k = 0; k < np_old; ++k) {
            double xk = ws.xs[k][tid];
            xk = fma(tmin, ws.bs[k][tid] - xk, xk);
            if (k == cand) xk = 0.0;
            if (xk > 0.0) {
                ws.xs[q][tid] = xk;
                ws.Ps[q][tid] = ws.Ps[k][tid];
                ++q;
            } else {
                rmask |= 1u << k;
                removed(ws.Ps[k][tid]);
            }
        }
        np = q;
        if (np == 0) return true;
        for (int pn = np_old; rmask; --pn) {  // Cholesky downdate, highest removed position first (chol_delete, per thread)
            const int qd = 31 - __clz(rmask);
            rmask &= ~(1u << qd);
            double carry[CAPT];
#pragma unroll
            for (int i = 0; i < CAPT; ++i) carry[i] = 0.0;
            for (int i = qd; i < pn - 1; ++i)
                for (int col = 0; col < qd; ++col) ws.Ls[tri(i, col)][tid] = ws.Ls[tri(i + 1, col)][tid];
#pragma unroll
            for (int i = 0; i < CAPT; ++i)
                if (i >= qd && i < pn - 1) carry[i] = ws.Ls[tri(i + 1, qd)][tid];
            for (int r = qd; r < pn - 1; ++r) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < CAPT; ++i)
                    if (i == r) a = carry[i];
                const double b = ws.Ls[tri(r + 1, r + 1)][tid];
                const double ir = rsqrt(fma(a, a, b * b));
                const double cs = a * ir, sn = b * ir;
#pragma unroll
                for (int i = 0; i < CAPT; ++i) {
                    if (i >= r && i < pn - 1) {
                        const double u2 = ws.Ls[tri(i + 1, r + 1)][tid];
                        ws.Ls[tri(i, r)][tid] = fma(cs, carry[i], sn * u2);
                        carry[i] = fma(cs, u2, -sn * carry[i]);
                    }
                }
                ws.rds[r][tid] = ir;
                const double zr = ws.zs[r][tid], zr1 = ws.zs[r + 1][tid];
                ws.zs[r][tid] = fma(cs, zr, sn * zr1);
                ws.zs[r + 1][tid] = fma(cs, zr1, -sn * zr);
            }
            ws.zs[pn - 1][tid] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NODDI stage 3 (debias on the support, amico/models.pyx:929-942 + maps :945-979): the system is the sub-system on the ~11
// support atoms (compact numbering, sa[] maps it to atoms), small enough for the dual pass to stay per thread as well.
template <int CAPT>
__host__ __device__ constexpr int tpv3_smem_bytes() { return TPV_THREADS * (TpvWS<CAPT>::BYTES_PER_THREAD + 32); }

template <int NPL, int CAPT>
__global__ void __launch_bounds__(TPV_THREADS, 4) k_noddi_stage3_tpv(const FitParams p, int4 *redo, int *redo_count)
{
    constexpr int NTH = TPV_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TpvWS<CAPT> ws;
    unsigned char (*sa)[NTH] = (unsigned char (*)[NTH])ws.carve(smem_raw);  // compact index -> atom
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_wm = p.n_wm, ldT = p.ldT1, NA = p.NA, m = p.m;
    const bool use_as = p.aspace != 0;
    for (long long base = (long long)blockIdx.x * NTH; base < p.n_vox; base += (long long)gridDim.x * NTH) {
        const long long pos = base + tid;
        const bool active = pos < p.n_vox;
        const long long vox = active ? (long long)p.order[pos] : 0;
        const int dir = active ? p.lut[vox] : 0;
        const double *T = p.T1 + (size_t)dir * p.T1_stride;
        const double *cg = p.c1_all + (size_t)(active ? pos : 0) * NA;
        asm volatile("" : "+l"(T), "+l"(cg));  // opaque: held in registers, not re-derived from the kernel parameters at every load
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
        auto atom_of = [&](int k) { return (int)sa[ws.Ps[k][tid]][tid]; };
        auto atom_of_thread = [&](int k, int st) { return (int)sa[ws.Ps[k][st]][st]; };
        while (__any_sync(FULL, run)) {
            bool pend = run && np < m;
            if (pend && np >= CAPT) { pend = false; run = false; redo_me = true; }
            if (run && !pend) run = false;  // np reached m: done
            unsigned valid = allmask & ~inP;
            bool acc = false;
            int j = -1, atomj = 0;
            double d2 = 0.0, znum = 0.0, vv = 0.0, hjj = 0.0;
            while (__any_sync(FULL, pend)) {
                bool near = false;
                if (pend) {
                    // dual w = c - T[:,P] x_P over the still valid atoms, largest positive one (ties: lowest index)
                    double bv = 0.0;
                    j = -1;
                    int prow[CAPT];     // row offsets of the passive atoms and their coefficients, in registers:
                    double xk[CAPT];    // the Gram entries of an atom are then loaded together, not one per dependent step
#pragma unroll
                    for (int k = 0; k < CAPT; ++k) {
                        const bool on = k < np;
                        prow[k] = on ? atom_of(k) * ldT : 0;
                        xk[k] = on ? ws.xs[k][tid] : 0.0;
                    }
                    for (unsigned mm = valid; mm;) {  // two atoms per trip: both atoms' Gram entries in flight together
                        const int a0 = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const bool two = mm != 0u;
                        const int a1 = two ? __ffs(mm) - 1 : a0;
                        mm &= mm - 1;
                        const int atom0 = sa[a0][tid], atom1 = sa[a1][tid];
                        const double *T0 = T + atom0, *T1p = T + atom1;
                        double w0 = cg[atom0], w1 = cg[atom1];
                        double g0[CAPT], g1[CAPT];
#pragma unroll
                        for (int k = 0; k < CAPT; ++k) {
                            g0[k] = (k < np) ? T0[prow[k]] : 0.0;
                            g1[k] = (k < np) ? T1p[prow[k]] : 0.0;
                        }
#pragma unroll
                        for (int k = 0; k < CAPT; ++k)
                            if (k < np) {
                                w0 = fma(-g0[k], xk[k], w0);
                                w1 = fma(-g1[k], xk[k], w1);
                            }
                        if (w0 > bv) { bv = w0; j = a0; }
                        if (two && w1 > bv) { bv = w1; j = a1; }
                    }
                    if (j < 0) { pend = false; run = false; }
                }
                if (pend) {
                    atomj = sa[j][tid];
                    const double *Tj = T + (unsigned)(atomj * ldT);
                    hjj = Tj[atomj];
                    tpv_candidate<CAPT>(ws, tid, np, Tj, hjj, cg[atomj], atom_of, vv, d2, znum);
                    near = use_as && np > 0 && d2 < 1e-10 * hjj;
                    if (near) tpv_back_inplace<CAPT>(ws, tid, np);  // beta for the service
                }
                tpv_aspace_service<CAPT>(ws, p, near, tid, lane, np, atomj, dir, vox, hjj, atom_of_thread, d2, znum);
                if (pend) {
                    if (near) tpv_restore_v<CAPT>(ws, tid, np, T + (unsigned)(atomj * ldT), atom_of);
                    if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) { acc = true; pend = false; }
                    else valid &= ~(1u << j);
                }
            }
            if (!acc) continue;
            inP |= 1u << j;
            if (!tpv_accept<CAPT>(ws, tid, np, iter, itmax, j, d2, znum, [&](int a) { inP &= ~(1u << a); })) run = false;
        }
        // ---- maps (noddi_maps: sums over the positive coefficients in ascending atom order)
        if (active && !redo_me) {
            auto xof = [&](int a) {  // coefficient of compact index a (passive atoms only)
                double xv = 0.0;
                for (int k = 0; k < np; ++k)
                    if (ws.Ps[k][tid] == a) xv = ws.xs[k][tid];
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
            if (p.coeff_out || (p.flags & (FLAG_RMSE | FLAG_NRMSE))) {
                // passive atoms in ascending atom order (<= CAPT of them)
                int sat[CAPT];
                double sxv[CAPT];
                unsigned mm = inP;
#pragma unroll
                for (int q = 0; q < CAPT; ++q) {
                    const bool on = mm != 0u;
                    const int a = on ? __ffs(mm) - 1 : 0;
                    mm &= mm - 1;
                    sat[q] = on ? (int)sa[a][tid] : 0;
                    sxv[q] = on ? xof(a) : 0.0;
                }
                if (p.coeff_out) {
                    double *co = p.coeff_out + vox * p.n;
                    for (int jj = 0; jj < p.n; ++jj) co[jj] = 0.0;
#pragma unroll
                    for (int q = 0; q < CAPT; ++q)
                        if (q < np) co[sat[q]] = sxv[q];
                }
                if (p.flags & (FLAG_RMSE | FLAG_NRMSE)) {  // fit_errors (amico/models.pyx:45-71), same operation order
                    const float *S = (const float *)p.slab + (size_t)dir * p.slab_stride;
                    const float *yf = (const float *)p.y + vox * m;
                    const double *yd = (const double *)p.y + vox * m;
                    double den = 0.0;
                    if (p.flags & FLAG_NRMSE)
                        for (int i = 0; i < m; ++i) {
                            const double yi = p.y_f64 ? yd[i] : (double)yf[i];
                            den = madd(den, yi, yi);
                        }
                    double acc_r = 0.0, acc_n = 0.0;
                    for (int i = 0; i < m; ++i) {
                        const float *Si = S + (size_t)i * p.n_pad;
                        double ye = 0.0;
#pragma unroll
                        for (int q = 0; q < CAPT; ++q)
                            if (q < np && sxv[q] != 0.0) ye = madd(ye, (double)Si[sat[q]], sxv[q]);
                        const double d = (p.y_f64 ? yd[i] : (double)yf[i]) - ye, dq = d * d;
                        acc_r += dq / (double)m;
                        if (den > 1e-16) acc_n += dq / den;
                    }
                    if (p.flags & FLAG_RMSE) p.rmse[vox] = sqrt(acc_r);
                    if (p.flags & FLAG_NRMSE) p.nrmse[vox] = den > 1e-16 ? sqrt(acc_n) : 0.0;
                }
            }
        }
        if (redo_me) {
            const int slot = atomicAdd(redo_count, 1);
            redo[slot] = make_int4(dir, (int)pos, 1, 0);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// NODDI stage 1 (isotropic fraction, amico/models.pyx:911) on the full dictionary.  A warp pulls four batches (<= 32 voxels)
// from the queue, computes their c1 = A^T y on the tensor pipe (gemm_c1, kept per voxel for stage 3), then lane 8 t + i owns
// voxel i of batch t.  Dual pass + arg-max are served cooperatively (see above); inP / rej: per-voxel bit words over the atoms
// (word s, bit l <-> atom l + 32 s) of the passive set and of the candidates dropped in the current round.
template <int NPL, int CAPT>
__host__ __device__ constexpr int tpv1_smem_bytes() { return TPV_THREADS * (TpvWS<CAPT>::BYTES_PER_THREAD + 8 * NPL) + (TPV_THREADS / 32) * BV * 8; }

template <int NPL, int CAPT>
__global__ void __launch_bounds__(TPV_THREADS) k_noddi_stage1_tpv(const FitParams p, int4 *redo, int *redo_count)
{
    constexpr int NTH = TPV_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TpvWS<CAPT> ws;
    unsigned (*inPw)[NTH] = (unsigned (*)[NTH])ws.carve(smem_raw);
    unsigned (*rejw)[NTH] = inPw + NPL;
    double *bx = (double *)(rejw + NPL) + (threadIdx.x >> 5) * BV;  // per warp: ||y||^2 of the batch in flight
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = p.n, ldT = p.ldT1, NA = p.NA, m = p.m;
    const bool use_as = p.aspace != 0;
    constexpr int NT = 4 * NPL, TP = (NT % 2 == 0) ? NT / 2 : NT;
    unsigned availw = 0u;  // bit s: atom lane + 32 s exists
#pragma unroll
    for (int s = 0; s < NPL; ++s) availw |= (lane + 32 * s < n ? 1u : 0u) << s;
    int *counter = p.tile_counter;
    const int n_tiles = *p.n_tiles_ptr;
    const int itmax = 3 * n;
    for (;;) {
        int b0 = 0;
        if (lane == 0) b0 = atomicAdd(counter, 4);
        b0 = __reduce_add_sync(FULL, b0);
        if (b0 >= n_tiles) break;
        bool active = false;
        long long pos = 0, vox = 0;
        int dir = 0;
        double yy = 0.0;
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
            if (b0 + t >= n_tiles) break;
            const int4 tile = p.tiles[b0 + t];
            const int g = lane >> 2;
            const bool vvalid = g < tile.z;
            const float *S = (const float *)p.slab + (size_t)tile.x * p.slab_stride;
            gemm_c1<NT, TP, float>(S, p.n_pad, m, p.y, p.y_f64, (long long)p.order[tile.y + (vvalid ? g : 0)], vvalid,
                                   p.c1_all + (size_t)tile.y * NA, NA, lane, bx);
            if ((lane >> 3) == t && (lane & 7) < tile.z) {
                active = true;
                pos = tile.y + (lane & 7);
                dir = tile.x;
                yy = bx[lane & 7];
            }
            __syncwarp();
        }
        if (active) vox = (long long)p.order[pos];
        const double *T = p.T1 + (size_t)dir * p.T1_stride;
        const double *cg = p.c1_all + (size_t)pos * NA;
        bool redo_me = false, run = active;
        int np = 0, iter = 0;
#pragma unroll
        for (int s = 0; s < NPL; ++s) inPw[s][tid] = 0u;
        auto atom_of = [&](int k) { return ws.Ps[k][tid]; };
        auto atom_of_thread = [&](int k, int st) { return ws.Ps[k][st]; };
        while (__any_sync(FULL, run)) {
            bool pend = run && np < m;
            if (pend && np >= CAPT) { pend = false; run = false; redo_me = true; }
            if (run && !pend) run = false;
#pragma unroll
            for (int s = 0; s < NPL; ++s) rejw[s][tid] = inPw[s][tid];  // blocked atoms of this round: passive + dropped candidates
            bool acc = false;
            int j = -1;
            double d2 = 0.0, znum = 0.0, vv = 0.0, hjj = 0.0;
            while (__any_sync(FULL, pend)) {
                __syncwarp();  // passive lists, coefficients and bit words of the requesters are read by the other lanes
                // ---- dual pass + arg-max service: w = c - T[:,P] x_P in passive order, lane l owns atoms l + 32 s
                for (unsigned req = __ballot_sync(FULL, pend); req; req &= req - 1) {
                    const int src = __ffs(req) - 1, st = (tid & ~31) + src;
                    const int r_np = __shfl_sync(FULL, np, src), r_dir = __shfl_sync(FULL, dir, src);
                    const long long r_pos = __shfl_sync(FULL, pos, src);
                    const double *Tl = p.T1 + (size_t)r_dir * p.T1_stride + lane;
                    const double *cl = p.c1_all + (size_t)r_pos * NA + lane;
                    // rows in chunks of four, every load of a chunk issued before its first use (one L2 latency per chunk)
                    double wl[NPL];
#pragma unroll
                    for (int s = 0; s < NPL; ++s) wl[s] = cl[32 * s];
#pragma unroll 1
                    for (int k0 = 0; k0 < r_np; k0 += 4) {
                        double gq[4][NPL], xq[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int k = min(k0 + q, r_np - 1);
                            const double *rk = Tl + (unsigned)(ws.Ps[k][st] * ldT);
                            xq[q] = ws.xs[k][st];
#pragma unroll
                            for (int s = 0; s < NPL; ++s) gq[q][s] = rk[32 * s];
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (k0 + q < r_np) {
#pragma unroll
                                for (int s = 0; s < NPL; ++s) wl[s] = fma(-gq[q][s], xq[q], wl[s]);
                            }
                        }
                    }
                    double bv = 0.0;
                    int bj = -1;
#pragma unroll
                    for (int s = 0; s < NPL; ++s) {
                        const bool ok = ((availw >> s) & 1u) && !((rejw[s][st] >> lane) & 1u);
                        if (ok && wl[s] > bv) { bv = wl[s]; bj = lane + 32 * s; }
                    }
                    warp_argmax_pos(bv, bj);
                    if (lane == src) j = bj;
                }
                if (pend && j < 0) { pend = false; run = false; }  // no positive dual left: done
                bool near = false;
                if (pend) {
                    const double *Tj = T + (unsigned)(j * ldT);
                    hjj = Tj[j];
                    tpv_candidate<CAPT>(ws, tid, np, Tj, hjj, cg[j], atom_of, vv, d2, znum);
                    near = use_as && np > 0 && d2 < 1e-10 * hjj;
                    if (near) tpv_back_inplace<CAPT>(ws, tid, np);
                }
                tpv_aspace_service<CAPT>(ws, p, near, tid, lane, np, j, dir, vox, hjj, atom_of_thread, d2, znum);
                if (pend) {
                    if (near) tpv_restore_v<CAPT>(ws, tid, np, T + (unsigned)(j * ldT), atom_of);
                    if (d2 > 0.0 && d2 > 1.2325951644078309e-28 * vv && znum > 0.0) { acc = true; pend = false; }
                    else rejw[j >> 5][tid] |= 1u << (j & 31);
                }
                __syncwarp();  // rej / passive words are read by the other lanes in the next service round
            }
            if (acc) {
                inPw[j >> 5][tid] |= 1u << (j & 31);
                if (!tpv_accept<CAPT>(ws, tid, np, iter, itmax, j, d2, znum, [&](int a) { inPw[a >> 5][tid] &= ~(1u << (a & 31)); })) run = false;
            }
            __syncwarp();
        }
        if (active && !redo_me) {
            double x1 = 0.0, x2 = 0.0, zz = 0.0;
            for (int k = 0; k < np; ++k) {
                if (ws.Ps[k][tid] == n - 1) x1 = ws.xs[k][tid];
                if (ws.Ps[k][tid] == n - 2) x2 = ws.xs[k][tid];
            }
            {   // ||z||^2 over the 32-lane butterfly tree, as warp_sum would add it
                double q[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double zk = (k < np && k < CAPT) ? ws.zs[k < CAPT ? k : 0][tid] : 0.0;
                    q[k] = zk * zk;
                }
                zz = ((q[0] + q[4]) + (q[2] + q[6])) + ((q[1] + q[5]) + (q[3] + q[7]));
            }
            p.xiso[2 * pos] = x1;
            p.xiso[2 * pos + 1] = p.exvivo ? x2 : 0.0;
            if (yy > 0.0 && yy - zz < p.exact_tol * yy) {  // exact-fit voxel: queued for the A-space QR path (see k_noddi_stage)
                const unsigned long long idx = atomicAdd((unsigned long long *)&p.status[4], 1ull);
                if ((long long)idx < p.exact_cap) p.exact_list[idx] = (int)vox;
            }
        }
        if (redo_me) {
            const int slot = atomicAdd(redo_count, 1);
            redo[slot] = make_int4(dir, (int)pos, 1, 0);
        }
        __syncwarp();
    }
}

}  // namespace amx
