import numpy as np, sys, time
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from amico_b200 import synth
from oracle import oracle as orc

def lh_gram_chol(H, c):
    """Gram-space LH; passive-set solves by Cholesky of H_PP recomputed each time (fp64)."""
    n=len(c); x=np.zeros(n); P=[]; inP=np.zeros(n,bool); it=0
    while len(P)<n:
        w = c - H[:,P]@x[P] if P else c.copy()
        w[inP]=0
        while True:
            j=int(np.argmax(np.where(inP,-np.inf,w)))
            if w[j]<=0: return x
            if P:
                L=np.linalg.cholesky(H[np.ix_(P,P)])
                v=solve_triangular(L,H[P,j],lower=True)
                d2=H[j,j]-v@v
                unorm=np.sqrt(v@v)
                if d2>0 and (unorm+np.sqrt(d2)*0.01)-unorm>0:
                    z=solve_triangular(L,c[P],lower=True)
                    zt=(c[j]-v@z)/d2
                    if zt>0: break
            else:
                if H[j,j]>0 and c[j]/H[j,j]>0: break
            w[j]=0
        P.append(j); inP[j]=True
        while True:
            it+=1
            if it>3*n: return x
            L=np.linalg.cholesky(H[np.ix_(P,P)])
            s=cho_solve((L,True),c[P])
            if (s>0).all(): x[:]=0; x[P]=s; break
            xp=x[P]; neg=s<=0
            t=np.where(neg, xp/(xp-s), np.inf); k=int(np.argmin(t)); al=t[k]
            xp=xp+al*(s-xp); xp[k]=0
            x[:]=0; x[P]=xp
            rem=[q for q in range(len(P)) if xp[q]<=0]
            for q in sorted(rem,reverse=True):
                inP[P[q]]=False; x[P[q]]=0; del P[q]
            if not P: break
    return x

if __name__=='__main__':
    n_vox=int(sys.argv[1])
    P=synth.make_problem(2,n_vox=n_vox); K=P.KERNELS
    lut=synth.lut_index_numpy(P.DIRs,P.htable); Hc={}
    bad=0; diffs=[]
    for i in range(n_vox):
        k=lut[i]
        if k not in Hc:
            A=synth.dictionary_for_direction('NODDI',K,k); Hc[k]=(A,A.T@A)
        A,H=Hc[k]; y=P.y[i].astype(np.float64)
        xo,_=orc.nnls(A,y); xg=lh_gram_chol(H,A.T@y)
        diffs.append(np.abs(xg-xo).max()); bad+=((xg>0)!=(xo>0)).any()
    diffs=np.array(diffs)
    print('mismatch',bad,'of',n_vox,'max',diffs.max(),'p99.9',np.percentile(diffs,99.9),'p99',np.percentile(diffs,99),'p50',np.median(diffs))
