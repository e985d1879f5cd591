import numpy as np, time, sys
from amico_b200 import synth
from oracle import oracle as orc

def lh_gram(H, c, mode='inv', tol_indep=1e-14):
    n = len(c); x = np.zeros(n); P = []; M = np.zeros((0,0))
    inP = np.zeros(n, bool)
    w = c.copy()
    it = 0
    while len(P) < n:
        w = c - H[:, P] @ x[P] if P else c.copy()
        w[inP] = 0
        accepted = False
        while True:
            j = int(np.argmax(np.where(inP, -np.inf, w)))
            if w[j] <= 0: return x, P
            # candidate test
            if P:
                g = H[P, j]
                u = M @ g
                d2 = H[j, j] - g @ u
                if not (d2 > tol_indep * H[j, j]):
                    w[j] = 0; continue
            else:
                d2 = H[j, j]
            # new coefficient
            if w[j] / d2 > 0: break
            w[j] = 0
        # add j
        if P:
            Mn = np.zeros((len(P)+1,)*2)
            Mn[:-1,:-1] = M + np.outer(u,u)/d2
            Mn[:-1,-1] = -u/d2; Mn[-1,:-1] = -u/d2; Mn[-1,-1] = 1/d2
            M = Mn
        else:
            M = np.array([[1/d2]])
        P.append(j); inP[j] = True
        while True:
            it += 1
            if it > 3*n: return x, P
            if mode == 'inv': s = M @ c[P]
            else: s = np.linalg.solve(H[np.ix_(P,P)], c[P])
            if (s > 0).all():
                x[:] = 0; x[P] = s; break
            xp = x[P]
            neg = s <= 0
            t = np.where(neg, -xp/(s-xp+ (~neg)*1.0), np.inf)
            t = np.where(neg, xp/(xp-s), np.inf)
            k = int(np.argmin(t)); alpha = t[k]
            xp = xp + alpha*(s-xp)
            xp[k] = 0
            rem = [k]
            # remove also any x<=0
            for q in range(len(P)):
                if q != k and xp[q] <= 0: rem.append(q)
            x[:] = 0; x[P] = xp
            for q in sorted(rem, reverse=True):
                # downdate inverse
                piv = M[q,q]; col = np.delete(M[:,q], q)
                M = np.delete(np.delete(M, q, 0), q, 1) - np.outer(col,col)/piv
                inP[P[q]] = False; x[P[q]] = 0
                del P[q]
            if not P: break
    return x, P

if __name__ == '__main__':
    n_vox = int(sys.argv[1]) if len(sys.argv)>1 else 1000
    P = synth.make_problem(2, n_vox=n_vox)
    K = P.KERNELS
    lut = synth.lut_index_numpy(P.DIRs, P.htable)
    Hc = {}
    bad = 0; md = 0; t0=time.time(); diffs=[]
    for i in range(n_vox):
        k = lut[i]
        if k not in Hc:
            A = synth.dictionary_for_direction('NODDI', K, k); Hc[k] = (A, A.T@A)
        A, H = Hc[k]
        y = P.y[i].astype(np.float64)
        xo, _ = orc.nnls(A, y)
        for mode in ('inv',):
            xg, _ = lh_gram(H, A.T@y, mode)
            d = np.abs(xg-xo).max(); diffs.append(d)
            if ((xg>0)!=(xo>0)).any(): bad += 1
    diffs=np.array(diffs)
    print('support mismatches', bad, 'of', n_vox, 'max diff', diffs.max(), 'p99', np.percentile(diffs,99), 'p50', np.median(diffs), time.time()-t0)
