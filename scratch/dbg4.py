import numpy as np, ctypes as C, sys
from amico_b200 import synth
from oracle import oracle as orc
lib = C.CDLL('/root/repo/scratch/libgm.so')
dp = C.POINTER(C.c_double)
def P_(a): return a.ctypes.data_as(dp)
n_vox=4000
P = synth.make_problem(2, n_vox=n_vox); K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
for i in range(n_vox):
    k=int(lut[i])
    A = np.asfortranarray(synth.dictionary_for_direction('NODDI', K, k)); H=np.ascontiguousarray(A.T@A)
    y = P.y[i].astype(np.float64); c = A.T@y
    n=A.shape[1]; m=A.shape[0]; x=np.zeros(n)
    lib.gm_nnls(P_(H), n, P_(c), n, m, P_(x), P_(A), P_(y), m, 0, None)
    xo,_ = orc.nnls(A,y)
    if ((x>0)!=(xo>0)).any():
        so=np.nonzero(xo>0)[0]; sg=np.nonzero(x>0)[0]
        print('vox',i,'\n oracle',so,xo[so],'\n gram  ',sg,x[sg])
        print(' resid oracle',np.linalg.norm(A@xo-y),'gram',np.linalg.norm(A@x-y), 'iso', xo[-1], x[-1])
        print(' sv A_S oracle', np.linalg.svd(A[:,so],compute_uv=False)[[0,-2,-1]], 'gram', np.linalg.svd(A[:,sg],compute_uv=False)[[0,-2,-1]])
        w=A.T@(y-A@xo); print(' oracle max dual', w.max(), 'gram max dual', (A.T@(y-A@x)).max())
