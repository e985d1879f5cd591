import numpy as np
from amico_b200 import synth
from oracle import oracle as orc
import scratch.proto_gram as pg
P = synth.make_problem(2, n_vox=300)
K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
i=2
A = synth.dictionary_for_direction('NODDI', K, lut[i]); H=A.T@A
y = P.y[i].astype(np.float64); c=A.T@y
# trace
n=len(c); x=np.zeros(n); Pset=[]
for it in range(30):
    w = c - H@x
    wz = w.copy(); wz[Pset]=-np.inf
    j=int(np.argmax(wz)); print('it',it,'P',Pset,'argmax',j,'w',w[j])
    if w[j]<=0: break
    if Pset:
        g=H[Pset,j]; u=np.linalg.solve(H[np.ix_(Pset,Pset)],g); d2=H[j,j]-g@u; print('   d2',d2,'Hjj',H[j,j], 'ratio', d2/H[j,j])
    Pset.append(j)
    while True:
        s=np.linalg.solve(H[np.ix_(Pset,Pset)],c[Pset])
        if (s>0).all(): x[:]=0; x[Pset]=s; break
        xp=x[Pset]; neg=s<=0
        t=np.where(neg, xp/(xp-s), np.inf); k=int(np.argmin(t)); al=t[k]
        xp=xp+al*(s-xp); print('   remove',Pset[k],'alpha',al)
        x[:]=0; x[Pset]=xp; x[Pset[k]]=0; del Pset[k]
print(np.nonzero(x)[0])
