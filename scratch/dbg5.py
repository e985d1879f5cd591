import numpy as np, sys
from scipy.linalg import solve_triangular
from amico_b200 import synth
from oracle import oracle as orc
P = synth.make_problem(2, n_vox=4000); K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
i=int(sys.argv[1])
A = synth.dictionary_for_direction('NODDI', K, int(lut[i])); H=A.T@A
y = P.y[i].astype(np.float64); c=A.T@y
n=len(c); x=np.zeros(n); Ps=[]; inP=np.zeros(n,bool)
for it in range(40):
    w = c - H[:,Ps]@x[Ps] if Ps else c.copy(); w[inP]=0
    acc=False
    while True:
        j=int(np.argmax(np.where(inP,-np.inf,w)))
        if w[j]<=0: print('done, wmax',w[j]); break
        if Ps:
            L=np.linalg.cholesky(H[np.ix_(Ps,Ps)]); v=solve_triangular(L,H[Ps,j],lower=True); z=solve_triangular(L,c[Ps],lower=True)
            d2=H[j,j]-v@v; zn=(c[j]-v@z)
            # A-space truth
            Q,R=np.linalg.qr(A[:,Ps]); aj=A[:,j]-Q@(Q.T@A[:,j]); aj=aj-Q@(Q.T@aj)
            print(f'  cand {j} w {w[j]:.3e} d2 {d2:.3e} true d2 {aj@aj:.3e} zn {zn:.3e} true zn {aj@y:.3e}')
            if d2>0 and zn>0: acc=True; break
        else:
            acc=True; break
        w[j]=0
    if not acc: break
    Ps.append(j); inP[j]=True
    while True:
        s=np.linalg.solve(H[np.ix_(Ps,Ps)],c[Ps])
        if (s>0).all(): x[:]=0; x[Ps]=s; break
        xp=x[Ps]; neg=s<=0
        t=np.where(neg, xp/(xp-s), np.inf); k=int(np.argmin(t)); al=t[k]
        xp=xp+al*(s-xp); xp[k]=0; x[:]=0; x[Ps]=xp
        rem=[q for q in range(len(Ps)) if xp[q]<=0]
        for q in sorted(rem,reverse=True):
            print('   remove',Ps[q],'alpha',al); inP[Ps[q]]=False; x[Ps[q]]=0; del Ps[q]
    print('it',it,'P',Ps)
