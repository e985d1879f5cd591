"""ctypes front-end of the CPU oracle (``oracle/amico_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, never by the ``amico_b200`` package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_i16_p = C.POINTER(C.c_int16)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "amico_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def load():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def nnls(A, y):
    """Lawson-Hanson NNLS; A (m, n).  Returns x, rnorm."""
    lib = load()
    A = np.asfortranarray(A, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m, n = A.shape
    x = np.zeros(n)
    rn = C.c_double(0.0)
    lib.orc_nnls(_p(A, c_double_p), _p(y, c_double_p), C.c_int(m), C.c_int(n), _p(x, c_double_p), C.byref(rn))
    return x, rn.value


def lasso(A, y, lambda1, lambda2):
    """SPAMS-style non-negative elastic net (LARS, PENALTY mode); A (m, n), y (m,)."""
    lib = load()
    A = np.asfortranarray(A, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m, n = A.shape
    x = np.zeros(n)
    lib.orc_lasso(_p(A, c_double_p), _p(y, c_double_p), C.c_int(m), C.c_int(n), C.c_int(1), _p(x, c_double_p),
                  C.c_double(lambda1), C.c_double(lambda2))
    return x


def lut_indices(dirs, htable):
    """LUT index per direction (int32; -1 = out of range).  ``dirs`` is NOT modified (a copy is flipped)."""
    lib = load()
    d = np.array(dirs, dtype=np.float64, order="C")
    ht = np.ascontiguousarray(htable, dtype=np.int16)
    idx = np.zeros(len(d), dtype=np.int32)
    lib.orc_lut_indices(_p(d, c_double_p), C.c_int64(len(d)), _p(ht, c_i16_p), _p(idx, c_i32_p))
    return idx


FLAG_RMSE, FLAG_NRMSE, FLAG_EXTRA = 1, 2, 4


def fit(model, y, dirs, htable, K, params, lambda1, lambda2, *, rmse=False, nrmse=False, extra=False,
        nthreads=1, dwi_idx=None, return_debug=False):
    """Run the oracle's per-voxel fit.  Mirrors the dict ``<Model>.fit`` returns (models.pyx:185-203).

    ``dirs`` is copied; pass ``return_debug=True`` to also get ``lut`` (and ``support`` for NODDI)
    and the flipped ``dirs``.
    """
    lib = load()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n_vox, m = y.shape
    flags = (FLAG_RMSE if rmse else 0) | (FLAG_NRMSE if nrmse else 0) | (FLAG_EXTRA if extra else 0)
    out_rmse = np.zeros(n_vox) if rmse else None
    out_nrmse = np.zeros(n_vox) if nrmse else None
    err = C.c_int(-1)
    res = {}
    d = None if dirs is None else np.array(dirs, dtype=np.float64, order="C")
    ht = None if htable is None else np.ascontiguousarray(htable, dtype=np.int16)
    lut = np.zeros(n_vox, dtype=np.int32)
    if model == "NODDI":
        wm = np.ascontiguousarray(K["wm"], dtype=np.float32)
        iso = np.ascontiguousarray(K["iso"], dtype=np.float32)
        norms = np.ascontiguousarray(K["norms"], dtype=np.float64)
        icvf = np.ascontiguousarray(K["icvf"], dtype=np.float32)
        kappa = np.ascontiguousarray(K["kappa"], dtype=np.float32)
        dwi = np.ascontiguousarray(dwi_idx, dtype=np.int64)
        exvivo = 1 if params.get("isExvivo") else 0
        est = np.zeros((n_vox, 4 if exvivo else 3))
        mod = np.zeros((n_vox, 2)) if extra else None
        sup = np.zeros(n_vox, dtype=np.int32)
        st = lib.orc_fit_noddi(_p(y, c_double_p), _p(d, c_double_p), C.c_int64(n_vox), C.c_int(m), _p(ht, c_i16_p),
                               C.c_int(wm.shape[1]), _p(wm, c_float_p), C.c_int(wm.shape[0]), _p(iso, c_float_p),
                               _p(norms, c_double_p), _p(icvf, c_float_p), _p(kappa, c_float_p), _p(dwi, c_i64_p),
                               C.c_int(len(dwi)), C.c_int(exvivo), C.c_double(lambda1), C.c_double(lambda2),
                               C.c_int(flags), C.c_int(nthreads), _p(est, c_double_p), _p(out_rmse, c_double_p),
                               _p(out_nrmse, c_double_p), _p(mod, c_double_p), _p(lut, c_i32_p), _p(sup, c_i32_p),
                               C.byref(err))
        if extra:
            res["estimates_mod"] = mod
        if return_debug:
            res["support"] = sup
    elif model in ("FreeWater", "FreeWaterMouse"):
        D = np.ascontiguousarray(K["D"], dtype=np.float32)
        CSF = np.ascontiguousarray(K["CSF"], dtype=np.float32)
        mouse = 1 if params.get("type") == "Mouse" else 0
        est = np.zeros((n_vox, 4 if mouse else 2))
        yc = np.zeros((n_vox, m)) if extra else None
        st = lib.orc_fit_freewater(_p(y, c_double_p), _p(d, c_double_p), C.c_int64(n_vox), C.c_int(m),
                                   _p(ht, c_i16_p), C.c_int(D.shape[1]), _p(D, c_float_p), C.c_int(D.shape[0]),
                                   _p(CSF, c_float_p), C.c_int(CSF.shape[0]), C.c_int(mouse), C.c_double(lambda1),
                                   C.c_double(lambda2), C.c_int(flags), C.c_int(nthreads), _p(est, c_double_p),
                                   _p(out_rmse, c_double_p), _p(out_nrmse, c_double_p), _p(yc, c_double_p),
                                   _p(lut, c_i32_p), C.byref(err))
        if extra:
            res["y_corrected"] = yc
    elif model == "CylinderZeppelinBall":
        wmr = np.ascontiguousarray(K["wmr"], dtype=np.float32)
        wmh = np.ascontiguousarray(K["wmh"], dtype=np.float32)
        iso = np.ascontiguousarray(K["iso"], dtype=np.float32)
        Rs = np.ascontiguousarray(params["Rs"], dtype=np.float64)
        est = np.zeros((n_vox, 3))
        st = lib.orc_fit_czb(_p(y, c_double_p), _p(d, c_double_p), C.c_int64(n_vox), C.c_int(m), _p(ht, c_i16_p),
                             C.c_int(wmr.shape[1]), _p(wmr, c_float_p), C.c_int(wmr.shape[0]), _p(wmh, c_float_p),
                             C.c_int(wmh.shape[0]), _p(iso, c_float_p), C.c_int(iso.shape[0]), _p(Rs, c_double_p),
                             C.c_double(lambda1), C.c_double(lambda2), C.c_int(flags), C.c_int(nthreads),
                             _p(est, c_double_p), _p(out_rmse, c_double_p), _p(out_nrmse, c_double_p),
                             _p(lut, c_i32_p), C.byref(err))
    elif model == "SANDI":
        sig = np.asfortranarray(K["signal"], dtype=np.float64)
        norms = np.ascontiguousarray(K["norms"], dtype=np.float64)
        Rs = np.ascontiguousarray(params["Rs"], dtype=np.float64)
        d_in = np.ascontiguousarray(params["d_in"], dtype=np.float64)
        d_isos = np.ascontiguousarray(params["d_isos"], dtype=np.float64)
        est = np.zeros((n_vox, 6))
        st = lib.orc_fit_sandi(_p(y, c_double_p), C.c_int64(n_vox), C.c_int(m), _p(sig, c_double_p),
                               _p(norms, c_double_p), _p(Rs, c_double_p), C.c_int(len(Rs)), _p(d_in, c_double_p),
                               C.c_int(len(d_in)), _p(d_isos, c_double_p), C.c_int(len(d_isos)), C.c_double(lambda1),
                               C.c_double(lambda2), C.c_int(flags), C.c_int(nthreads), _p(est, c_double_p),
                               _p(out_rmse, c_double_p), _p(out_nrmse, c_double_p))
    else:
        raise ValueError(model)
    if st != 0:
        raise RuntimeError(f'"amico.lut.dir_to_lut_idx" index out of bounds (voxel {err.value})')
    res["estimates"] = est
    if rmse:
        res["rmse"] = out_rmse
    if nrmse:
        res["nrmse"] = out_nrmse
    if return_debug:
        res["lut"] = lut
        res["dirs"] = d
    return res


DEFAULT_LAMBDAS = {
    # set_solver defaults: models.pyx:721, :1077, :439, :1405
    "NODDI": (5e-1, 1e-3),
    "FreeWater": (0.0, 1e-3),
    "FreeWaterMouse": (0.0, 1e-3),
    "CylinderZeppelinBall": (0.0, 4.0),
    "SANDI": (0.0, 5e-3),
}


def fit_problem(P, **kw):
    """Oracle fit of a ``amico_b200.synth.Problem`` with the model's default lambdas."""
    l1, l2 = DEFAULT_LAMBDAS[P.model]
    l1 = kw.pop("lambda1", l1)
    l2 = kw.pop("lambda2", l2)
    return fit(P.model, P.y, P.DIRs, P.htable, P.KERNELS, P.params, l1, l2, dwi_idx=P.scheme.dwi_idx, **kw)
