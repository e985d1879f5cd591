import numpy as np, sys
from amico_b200 import synth
from oracle import oracle as orc
LD=np.longdouble
def chol(Hs):
    n=len(Hs); L=np.zeros((n,n),dtype=Hs.dtype)
    for i in range(n):
        for j in range(i+1):
            s=Hs[i,j]-L[i,:j]@L[j,:j]
            L[i,j]=np.sqrt(s) if i==j else s/L[j,j]
    return L
def fsub(L,b):
    z=np.zeros(len(b),dtype=L.dtype)
    for i in range(len(b)): z[i]=(b[i]-L[i,:i]@z[:i])/L[i,i]
    return z
def bsub(L,z):
    n=len(z); s=np.zeros(n,dtype=L.dtype)
    for i in range(n-1,-1,-1): s[i]=(z[i]-L[i+1:,i]@s[i+1:])/L[i,i]
    return s
def lh(H,c,dt):
    H=H.astype(dt); c=c.astype(dt)
    n=len(c); x=np.zeros(n,dtype=dt); P=[]; inP=np.zeros(n,bool); it=0
    while len(P)<n:
        w = c - H[:,P]@x[P] if P else c.copy()
        w[inP]=0
        while True:
            j=int(np.argmax(np.where(inP,-np.inf,w)))
            if w[j]<=0: return x
            if P:
                L=chol(H[np.ix_(P,P)]); v=fsub(L,H[P,j]); d2=H[j,j]-v@v; unorm=np.sqrt(v@v)
                if d2>0 and (unorm+np.sqrt(d2)*dt(0.01))-unorm>0:
                    z=fsub(L,c[P]); zt=(c[j]-v@z)/d2
                    if zt>0: break
            else:
                if c[j]/H[j,j]>0: break
            w[j]=0
        P.append(j); inP[j]=True
        while True:
            it+=1
            if it>3*n: return x
            L=chol(H[np.ix_(P,P)]); s=bsub(L,fsub(L,c[P]))
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
n_vox=int(sys.argv[1])
P=synth.make_problem(2,n_vox=n_vox); K=P.KERNELS
lut=synth.lut_index_numpy(P.DIRs,P.htable)
badd=[];badl=[]
for i in range(n_vox):
    A=synth.dictionary_for_direction('NODDI',K,lut[i]); y=P.y[i].astype(np.float64)
    xo,_=orc.nnls(A,y)
    Al=A.astype(LD); Hl=Al.T@Al; cl=Al.T@y.astype(LD)
    xd=lh((A.T@A),A.T@y,np.float64); xl=lh(Hl,cl,LD)
    if ((xd>0)!=(xo>0)).any(): badd.append(i)
    if ((xl>0)!=(xo>0)).any(): badl.append(i)
print('f64 mismatches',badd); print('f80 mismatches',badl)
