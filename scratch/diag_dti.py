import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from amico_b200 import _lib as L, synth
from amico_b200.evaluation import dti_design_matrix
lib = L.load()
sch = synth.make_scheme(2)
n = 1 << 20
y = torch.rand((n, 100), dtype=torch.float32, device='cuda') + 0.1
dirs = torch.empty((n, 3), dtype=torch.float64, device='cuda')
W = np.ascontiguousarray(np.linalg.pinv(dti_design_matrix(sch.b, sch.raw[:, :3]))[:6])
st = torch.cuda.current_stream().cuda_stream
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.amx_dti_directions(0, L.SPACE_DEVICE, y.data_ptr(), L.F32, n, 100, W.ctypes.data, 1e-4, dirs.data_ptr(), st)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(i, rc, 'call %.3f ms  sync %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter()
for i in range(5):
    lib.amx_dti_directions(0, L.SPACE_DEVICE, y.data_ptr(), L.F32, n, 100, W.ctypes.data, 1e-4, dirs.data_ptr(), st)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('5 back-to-back: enqueue %.3f ms, drain %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
