import numpy as np, ctypes as C, sys, time
from amico_b200 import synth
lib = C.CDLL(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'libgm.so'))
dp = C.POINTER(C.c_double)
def P_(a): return a.ctypes.data_as(dp)
n_vox = int(sys.argv[1])
P = synth.make_problem(2, n_vox=n_vox); K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
cache={}; tot=np.zeros(16,dtype=np.int32); per=[]
for i in range(n_vox):
    k=int(lut[i])
    if k not in cache:
        A = np.asfortranarray(synth.dictionary_for_direction('NODDI', K, k)); cache[k]=(A, np.ascontiguousarray(A.T@A))
    A,H = cache[k]; y = P.y[i].astype(np.float64); c = A.T@y
    n=A.shape[1]; m=A.shape[0]; x=np.zeros(n); st=np.zeros(16,dtype=np.int32)
    lib.gm_nnls(P_(H), n, P_(c), n, m, P_(x), P_(A), P_(y), m, 0, st.ctypes.data_as(C.POINTER(C.c_int)))
    tot+=st; per.append(st.copy())
per=np.array(per)
print('avg outer',tot[0]/n_vox,'iter',tot[1]/n_vox,'removals',tot[2]/n_vox,'cands',tot[3]/n_vox)
print('cands with d2rel <1e-6,1e-8,1e-10,1e-12 per voxel', tot[4:8]/n_vox)
print('voxels with any cand <1e-6,1e-8,1e-10,1e-12', (per[:,4:8]>0).mean(0))
print('max outer', per[:,0].max(), 'p99', np.percentile(per[:,0],99))
