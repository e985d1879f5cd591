import numpy as np, ctypes as C, sys, time
from amico_b200 import synth
from oracle import oracle as orc
lib = C.CDLL(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'libgm.so'))
dp = C.POINTER(C.c_double)
def P_(a): return a.ctypes.data_as(dp)
def maps_noddi(x, K, n_wm):
    s = x.sum() + 1e-16
    swm = (x[:n_wm]/s).sum() + 1e-16
    icvf = K['icvf']; kap = K['kappa']
    f1 = (icvf.astype(np.float64)*x[:n_wm]/s/swm).sum()
    f2 = ((1.0-icvf.astype(np.float64)).astype(np.float32).astype(np.float64)*x[:n_wm]/s/swm).sum()
    k1 = (kap.astype(np.float64)*x[:n_wm]/s/swm).sum()
    return np.array([f1/(f1+f2+1e-16), 2/np.pi*np.arctan2(1.0,k1), x[-1]/s])
import os
n_vox = int(sys.argv[1]); mode = int(sys.argv[2]); cfg = int(sys.argv[3]) if len(sys.argv)>3 else 2
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 30.0
C.c_double.in_dll(lib, 'gm_thr').value = float(os.environ.get('GM_THR', '1e-10'))
C.c_double.in_dll(lib, 'gm_floor').value = float(os.environ.get('GM_FLOOR', '1e-24'))
C.c_int.in_dll(lib, 'gm_lars_incr').value = int(os.environ.get('GM_LARS_INCR', '0'))
P = synth.make_problem(cfg, n_vox=n_vox, snr=snr, seed=104); K = P.KERNELS
ref = orc.fit_problem(P, return_debug=True, nthreads=8)
lut = ref['lut']
n_wm = K['wm'].shape[0]; n = n_wm+1; m = P.y.shape[1]
dwi = np.ascontiguousarray(P.scheme.dwi_idx, dtype=np.int64); dc = len(dwi)
norms = np.ascontiguousarray(K['norms']); iso = K['iso'].astype(np.float64)
cache = {}
est = np.zeros((n_vox,3)); sup = np.zeros(n_vox, dtype=np.int32)
x = np.zeros(n); s_ = C.c_int(0)
t0=time.time()
for i in range(n_vox):
    k = int(lut[i])
    if k not in cache:
        A = np.asfortranarray(synth.dictionary_for_direction('NODDI', K, k))
        A2 = A[dwi][:, :n_wm]*norms
        cache[k] = (A, np.ascontiguousarray(A.T@A), np.ascontiguousarray(A2.T@A2))
    A, T1, T2 = cache[k]
    y = P.y[i].astype(np.float64)
    lib.gm_noddi_voxel(P_(A), P_(T1), P_(T2), P_(y), m, n_wm, 0, dwi.ctypes.data_as(C.POINTER(C.c_int64)), dc, P_(norms), P_(iso),
                       C.c_double(0.5), C.c_double(1e-3), mode, P_(x), C.byref(s_))
    est[i] = maps_noddi(x, K, n_wm); sup[i] = s_.value
r = ref['estimates']
rel = np.abs(est-r)/np.maximum(np.abs(r),1e-3)
ok = (rel<=1e-4).all(1)
print('mode',mode,'n',n_vox,'pass frac',ok.mean(),'fails',(~ok).sum(),'support eq',(sup==ref['support']).mean(),'p50',np.median(rel),'p99',np.percentile(rel,99),'max',rel.max(), 'time',time.time()-t0)
print('failing idx', np.nonzero(~ok)[0][:20], 'lars sign flips', C.c_long.in_dll(lib, 'gm_lars_signflips').value,
      'nnls calls', C.c_long.in_dll(lib, 'gm_nnls_calls').value, 'with a near-dependent atom', C.c_long.in_dll(lib, 'gm_ill_nnls').value,
      'refined passive solves', C.c_long.in_dll(lib, 'gm_ill_solves').value)
