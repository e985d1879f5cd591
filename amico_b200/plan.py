"""Device-resident fit plan: KERNELS / hash table / Gram tables in HBM + the fused fit call.

Thin host wrapper over the C ABI (``include/amico_b200.h``).  ``Plan.fit`` accepts either host
numpy arrays (staged by the library) or CUDA torch tensors (zero-copy: their ``data_ptr`` is handed
to the kernels and the work is enqueued on torch's current stream).
"""
from __future__ import annotations

import ctypes as C

import os

import numpy as np

from . import _lib as L

_MODEL_IDS = {"NODDI": L.MODEL_NODDI, "FreeWater": L.MODEL_FREEWATER, "CylinderZeppelinBall": L.MODEL_CZB,
              "SANDI": L.MODEL_SANDI}


def _np(a, dtype, order="C"):
    a = np.asarray(a)
    if order == "F":
        return np.asfortranarray(a, dtype=dtype)
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Plan:
    """One model's tables on one GPU.

    Parameters mirror what ``<Model>._fit`` reads (amico/models.pyx:816-847, 1168-1190, 546-572, 1509-1530):
    ``KERNELS`` is the dict ``<Model>.resample`` returns, ``htable`` the int16 hash table of
    ``amico/lut.pyx:71-91``, ``params`` the model's physical grids, ``dwi_idx`` = ``scheme.dwi_idx``.
    """

    def __init__(self, model, KERNELS, htable=None, params=None, dwi_idx=None, device=0):
        lib = L.load()
        params = dict(params or {})
        self.model = model
        self.device = int(device)
        self._h = C.c_void_p()
        self._keep = []
        K = KERNELS
        if model not in _MODEL_IDS:
            raise ValueError(f"unknown model {model!r}")
        if model == "SANDI":
            sig = _np(K["signal"], np.float64, "F")
            norms = _np(K["norms"], np.float64)
            Rs, d_in, d_isos = (_np(params[k], np.float64).ravel() for k in ("Rs", "d_in", "d_isos"))
            m, n = sig.shape
            if n != len(Rs) + len(d_in) + len(d_isos) or len(norms) != n:
                raise ValueError("SANDI KERNELS do not match the model grids")
            L.check(lib.amx_plan_create_sandi(self.device, m, len(Rs), len(d_in), len(d_isos), _ptr(sig), _ptr(norms),
                                              _ptr(Rs), _ptr(d_in), _ptr(d_isos), C.byref(self._h)))
        else:
            if htable is None:
                raise ValueError("htable is required")
            ht = _np(htable, np.int16).ravel()
            if ht.size != 181 * 181:
                raise ValueError("htable must have 181*181 entries")
            if model == "NODDI":
                wm = _np(K["wm"], np.float32)
                n_wm, ndirs, m = wm.shape
                iso = _np(K["iso"], np.float32).ravel()
                norms = _np(K["norms"], np.float64)
                icvf, kappa = _np(K["icvf"], np.float32), _np(K["kappa"], np.float32)
                if dwi_idx is None:
                    raise ValueError("dwi_idx (scheme.dwi_idx) is required for NODDI")
                dwi = _np(dwi_idx, np.int64).ravel()
                if norms.shape != (len(dwi), n_wm) or iso.size != m or icvf.size != n_wm or kappa.size != n_wm:
                    raise ValueError("NODDI KERNELS have inconsistent shapes")
                L.check(lib.amx_plan_create_noddi(self.device, m, ndirs, n_wm, _ptr(wm), _ptr(iso), _ptr(norms), _ptr(icvf),
                                                  _ptr(kappa), _ptr(dwi), len(dwi), 1 if params.get("isExvivo") else 0,
                                                  _ptr(ht), C.byref(self._h)))
            elif model == "FreeWater":
                D = _np(K["D"], np.float32)
                CSF = _np(K["CSF"], np.float32)
                CSF = CSF.reshape(-1, D.shape[2])
                n_perp, ndirs, m = D.shape
                L.check(lib.amx_plan_create_freewater(self.device, m, ndirs, n_perp, _ptr(D), CSF.shape[0], _ptr(CSF),
                                                      1 if params.get("type") == "Mouse" else 0, _ptr(ht), C.byref(self._h)))
            else:
                wmr = _np(K["wmr"], np.float32)
                wmh = _np(K["wmh"], np.float32)
                n_rs, ndirs, m = wmr.shape
                wmh = wmh.reshape(-1, ndirs, m)
                iso = _np(K["iso"], np.float32).reshape(-1, m)
                Rs = _np(params["Rs"], np.float64).ravel()
                if len(Rs) != n_rs:
                    raise ValueError("len(Rs) does not match KERNELS['wmr']")
                L.check(lib.amx_plan_create_czb(self.device, m, ndirs, n_rs, _ptr(wmr), wmh.shape[0], _ptr(wmh), iso.shape[0],
                                                _ptr(iso), _ptr(Rs), _ptr(ht), C.byref(self._h)))
        info = [C.c_int() for _ in range(6)]
        L.check(lib.amx_plan_info(self._h, *[C.byref(i) for i in info]))
        _, self.m, self.n_atoms, self.n_maps, self.ndirs, _ = [i.value for i in info]

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.load().amx_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ fit
    def fit(self, y, dirs, lambda1, lambda2, *, rmse=False, nrmse=False, extra=False, debug=False, out=None, exact=False):
        """Fit every row of ``y``.

        Host path: ``y`` (n_vox, m) float32/float64 ndarray, ``dirs`` (n_vox, 3) float64 C-contiguous ndarray
        (FLIPPED IN PLACE like the reference) or None for SANDI.  Device path: the same as CUDA torch tensors.
        Returns a dict like ``<Model>.fit`` (amico/models.pyx:185-203).  ``exact=True`` (``AMX_FLAG_EXACT``): the bit-reproducible
        kernels that follow the reference's CPU arithmetic operation for operation (same maps to ~1e-12, lower throughput).
        """
        lib = L.load()
        flags = (L.FLAG_RMSE if rmse else 0) | (L.FLAG_NRMSE if nrmse else 0)
        has_extra = extra and self.model in ("NODDI", "FreeWater")
        if has_extra:
            flags |= L.FLAG_EXTRA
        if exact:
            flags |= L.FLAG_EXACT
        dev = _is_torch(y)
        a = L.FitArgs()
        a.lambda1, a.lambda2, a.flags = float(lambda1), float(lambda2), flags
        n_vox = int(y.shape[0])
        if y.ndim != 2 or y.shape[1] != self.m:
            raise ValueError(f"y must be (n_vox, {self.m})")
        a.n_vox = n_vox
        res = {}
        if dev:
            import torch
            if not y.is_cuda or y.device.index != self.device or not y.is_contiguous():
                raise ValueError("y must be a contiguous CUDA tensor on the plan's device")
            if y.dtype not in (torch.float32, torch.float64):
                raise ValueError("y must be float32 or float64")
            a.space, a.y_dtype, a.y = L.SPACE_DEVICE, (L.F64 if y.dtype == torch.float64 else L.F32), y.data_ptr()
            if self.model != "SANDI":
                if dirs is None or dirs.dtype != torch.float64 or not dirs.is_contiguous() or tuple(dirs.shape) != (n_vox, 3):
                    raise ValueError("dirs must be a contiguous float64 CUDA tensor (n_vox, 3)")
                a.dirs = dirs.data_ptr()
            mk = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype, device=y.device)
            a.stream = torch.cuda.current_stream(y.device).cuda_stream
            ptr = lambda t: t.data_ptr()
            i32 = torch.int32
        else:
            y = np.ascontiguousarray(y)
            if y.dtype not in (np.float32, np.float64):
                y = y.astype(np.float64)
            a.space, a.y_dtype, a.y = L.SPACE_HOST, (L.F64 if y.dtype == np.float64 else L.F32), y.ctypes.data
            if self.model != "SANDI":
                if not (isinstance(dirs, np.ndarray) and dirs.dtype == np.float64 and dirs.flags.c_contiguous
                        and dirs.shape == (n_vox, 3)):
                    raise ValueError("dirs must be a C-contiguous float64 ndarray (n_vox, 3)")
                a.dirs = dirs.ctypes.data
            mk = lambda *shape, dtype=np.float64: np.zeros(shape, dtype=dtype)
            ptr = lambda t: t.ctypes.data
            i32 = np.int32
        est = out if out is not None else mk(n_vox, self.n_maps)
        a.estimates = ptr(est)
        res["estimates"] = est
        if rmse:
            res["rmse"] = mk(n_vox)
            a.rmse = ptr(res["rmse"])
        if nrmse:
            res["nrmse"] = mk(n_vox)
            a.nrmse = ptr(res["nrmse"])
        if has_extra:
            key = "estimates_mod" if self.model == "NODDI" else "y_corrected"
            res[key] = mk(n_vox, 2 if self.model == "NODDI" else self.m)
            a.extra = ptr(res[key])
        if debug:
            res["lut"] = mk(n_vox, dtype=i32)
            res["support"] = mk(n_vox, dtype=i32)
            res["x"] = mk(n_vox, self.n_atoms)
            a.lut_out, a.support_out, a.coeff_out = ptr(res["lut"]), ptr(res["support"]), ptr(res["x"])
        err = C.c_int64(-1)
        self._keep = [y, dirs, res]
        rc = lib.amx_fit(self._h, C.byref(a), C.byref(err))
        if rc == L.AMX_E_LUT_RANGE:
            # same exception type and text as amico/lut.pyx:352-354
            raise RuntimeError(f'"amico.lut.dir_to_lut_idx" index out of bounds (voxel {err.value})')
        L.check(rc)
        return res

    def lut_indices(self, dirs):
        """LUT index of each direction; ``dirs`` float64 (n, 3) ndarray, flipped in place."""
        lib = L.load()
        if not (isinstance(dirs, np.ndarray) and dirs.dtype == np.float64 and dirs.flags.c_contiguous):
            raise ValueError("dirs must be a C-contiguous float64 ndarray")
        idx = np.zeros(len(dirs), dtype=np.int32)
        rc = lib.amx_lut_indices(self._h, L.SPACE_HOST, dirs.ctypes.data, len(dirs), idx.ctypes.data)
        if rc not in (L.AMX_OK, L.AMX_E_LUT_RANGE):
            L.check(rc)
        return idx

    def last_timing(self):
        """ms: (LUT index + binning, fused fit kernel, whole call on the device)."""
        out = (C.c_double * 8)()
        L.check(L.load().amx_plan_last_timing(self._h, out, 8))
        return {"binning_ms": out[0], "fit_kernel_ms": out[1], "total_ms": out[2]}

    def last_counters(self):
        out = (C.c_int64 * 16)()
        L.check(L.load().amx_plan_last_counters(self._h, out, 16))
        return {"launches": out[0], "tiles": out[1], "overflow_voxels": out[2], "smem_bytes": out[3], "warps_per_cta": out[4],
                "tma_staged": bool(out[5]), "slow_path_voxels": out[6], "grid": out[7], "exact_path_voxels": out[8]}
