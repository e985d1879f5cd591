import numpy as np, ctypes as C, sys, time
from amico_b200 import synth
from oracle import oracle as orc
lib = C.CDLL(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'libgm.so'))
dp = C.POINTER(C.c_double)
def P_(a): return a.ctypes.data_as(dp)
n_vox = int(sys.argv[1]); mode = int(sys.argv[2]); snr = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
P = synth.make_problem(2, n_vox=n_vox, snr=snr, seed=104); K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
import os
C.c_double.in_dll(lib, 'gm_thr').value = float(os.environ.get('GM_THR', '1e-6'))
C.c_double.in_dll(lib, 'gm_floor').value = float(os.environ.get('GM_FLOOR', '0'))
cache={}; bad=0; diffs=[]
for i in range(n_vox):
    k=int(lut[i])
    if k not in cache:
        A = np.asfortranarray(synth.dictionary_for_direction('NODDI', K, k)); cache[k]=(A, np.ascontiguousarray(A.T@A))
    A,H = cache[k]; y = P.y[i].astype(np.float64); c = A.T@y
    n=A.shape[1]; m=A.shape[0]; x=np.zeros(n)
    lib.gm_nnls(P_(H), n, P_(c), n, m, P_(x), P_(A), P_(y), m, mode, None)
    xo,_ = orc.nnls(A,y)
    d=np.abs(x-xo).max(); diffs.append(d); bad += ((x>0)!=(xo>0)).any()
diffs=np.array(diffs)
print('snr',snr,'mode',mode,'mismatch',bad,'of',n_vox,'max',diffs.max(),'p99',np.percentile(diffs,99),'p50',np.median(diffs))
