"""Drop-in model plugins: the ``amico.models`` classes' public surface with the fit on the B200.

Each class mirrors its reference counterpart (``amico/models.pyx``): same ``id`` / ``name`` /
``maps_name`` / ``maps_descr``, ``set`` / ``get_params`` / ``set_solver`` with the same arguments and
defaults, and ``fit(evaluation)`` reading the same ``evaluation`` attributes (``y``, ``DIRs``, ``htable``,
``KERNELS``, ``get_config``) and returning the same dict (``models.pyx:185-203``).  ``fit`` runs entirely
on the GPU through the C ABI (no chunking over host threads: ``evaluation.nthreads`` is ignored).

They can be injected into the reference without editing it through its own plugin hook
(``$AMICO_WIP_MODELS`` -> ``from amicowipmodels import *``, ``amico/models.pyx:20-26``; lookup by name in
``amico/core.py:290-291``) -- see INTEGRATION.md.

``generate`` writes the rotated SH-space atoms ``A_###.npy`` in the reference's on-disk layout (signal models:
``amico_b200.signals``, checked against the reference's ``synthesis.py``); ``resample`` reads such
files -- the reference's or ours -- and builds the same ``KERNELS`` dict as the reference, with the SH -> signal-space
projection on the GPU (``amx_resample_kernels``).
"""
from __future__ import annotations

import os

import numpy as np

from . import lut as _lut
from . import signals as _sig
from .plan import Plan

__all__ = ["NODDI", "FreeWater", "CylinderZeppelinBall", "SANDI", "BaseModel"]


def _fingerprint(a):
    """Cheap content fingerprint of a table for the plan cache: shape, dtype and a CRC of a strided sample (<= 64 Ki elements) plus
    the array's ends.  Object identity is not enough -- an id can be reused after garbage collection and KERNELS can be rebuilt in
    place -- and hashing 30-80 MB per ``fit`` call would cost more than the fit.  (An edit confined to elements the sample
    misses goes unnoticed: call ``invalidate_plan()`` after poking single entries into a table.)"""
    if a is None:
        return None
    import zlib
    a = np.asarray(a)
    flat = a.reshape(-1) if a.flags.c_contiguous else np.asfortranarray(a).reshape(-1, order="F") if a.flags.f_contiguous else a.ravel()
    step = max(1, flat.size // 65536)
    sample = np.ascontiguousarray(flat[::step])
    return (a.shape, a.dtype.str, zlib.crc32(sample.tobytes()), zlib.crc32(np.ascontiguousarray(flat[-64:]).tobytes()))


class BaseModel:
    """Common part of the plugin surface (``amico/models.pyx:75-217``)."""

    id = "BaseModel"
    name = "Base Model"

    def __init__(self):
        self.maps_name = []
        self.maps_descr = []
        self.scheme = None
        self._plan = None
        self._plan_key = None
        self.device = 0

    def set_solver(self):
        self.solver_params = {}

    def _atoms(self):
        """The model's atoms (``amico_b200.signals``) in the order the reference's ``generate`` writes them."""
        raise NotImplementedError

    def generate(self, out_path, aux, idx_in, idx_out, ndirs):
        """``<Model>.generate`` (``amico/models.pyx:428-480, 725-752, 1088-1111, 1409-1444``): write the rotated SH-space atoms
        ``A_###.npy`` to ``out_path``.  ``aux`` = ``amico_b200.lut.precompute_rotation_matrices(lmax, lut_directions)``."""
        for i, atom in enumerate(self._atoms()):
            np.save(os.path.join(out_path, f"A_{i + 1:03d}.npy"), _lut.rotate_kernel(atom, aux, idx_out, ndirs))

    def resample(self, in_path, idx_out, Ylm_out, doMergeB0, ndirs):
        raise NotImplementedError

    # -- shared pieces of the four resample() bodies -------------------------------------------------
    def _merge_idx(self, doMergeB0):
        """(nS, merge_idx) of ``amico/models.pyx:756-761`` (identical in every model)."""
        if doMergeB0:
            return 1 + self.scheme.dwi_count, np.hstack((self.scheme.b0_idx[0], self.scheme.dwi_idx))
        return self.scheme.nS, np.arange(self.scheme.nS)

    def _resample_atoms(self, in_path, first, count, isotropic, idx_out, Ylm_out, merge_idx, ndirs):
        """Atoms ``A_{first+1:03d}.npy`` ... of ``generate``'s output, projected to the scheme: float32
        (count, ndirs, len(merge_idx)), or (count, len(merge_idx)) for isotropic atoms."""
        if count == 0:
            return np.zeros((0, ndirs, len(merge_idx)) if not isotropic else (0, len(merge_idx)), dtype=np.float32)
        lms = []
        for i in range(count):
            lm = np.load(os.path.join(in_path, f"A_{first + i + 1:03d}.npy"))
            if not isotropic and lm.shape[0] != ndirs:
                raise RuntimeError('Outdated LUT. Call "generate_kernels( regenerate=True )" to update the LUT')
            lms.append(lm)
        return _lut.resample_kernels(np.stack(lms), self.scheme.nS, idx_out, Ylm_out, merge_idx, device=self.device)

    # -- plan cache: one upload + Gram precompute per (KERNELS, htable) -------------------------
    def _model_params(self):
        return {}

    def _get_plan(self, evaluation):
        K = evaluation.KERNELS
        if K.get("model") != self.id:
            raise RuntimeError("Response functions were not created with the same model")  # core.py:417-418
        ht = getattr(evaluation, "htable", None)
        dwi = getattr(self.scheme, "dwi_idx", None) if self.scheme is not None else None
        key = (tuple((k, _fingerprint(v)) for k, v in sorted(K.items()) if k != "model"), _fingerprint(ht), _fingerprint(dwi), self.device,
               repr(sorted((k, np.asarray(v).tobytes()) for k, v in self._model_params().items() if not isinstance(v, str))),
               tuple(sorted((k, v) for k, v in self._model_params().items() if isinstance(v, str))))
        if self._plan is None or self._plan_key != key:
            if self._plan is not None:
                self._plan.close()
            dwi_idx = getattr(self.scheme, "dwi_idx", None) if self.scheme is not None else None
            self._plan = Plan(self.id, K, ht, self._model_params(), dwi_idx=dwi_idx, device=self.device)
            self._plan_key = key
        return self._plan

    def invalidate_plan(self):
        """Drop the device-resident tables; the next ``fit`` uploads ``evaluation.KERNELS`` again."""
        if self._plan is not None:
            self._plan.close()
        self._plan = self._plan_key = None

    def fit(self, evaluation):
        """``<Model>.fit(evaluation)`` (``amico/models.pyx:795-811`` etc.) on the GPU."""
        if not hasattr(self, "solver_params"):
            self.set_solver()
        plan = self._get_plan(evaluation)
        y = evaluation.y
        y = np.ascontiguousarray(y) if y.dtype in (np.float32, np.float64) else np.ascontiguousarray(y, dtype=np.float64)
        dirs = None
        if self.id != "SANDI":
            src = evaluation.DIRs
            # The reference flips the hemisphere on np.ascontiguousarray(DIRs, dtype=double): a view -- hence
            # written through to evaluation.DIRs -- exactly when DIRs already is C-contiguous float64
            # (amico/lut.pyx:335-338 via models.pyx:835).  Reproduce that.
            dirs = np.ascontiguousarray(src, dtype=np.float64)
            if dirs.ndim != 2 or dirs.shape != (y.shape[0], 3):
                raise ValueError("evaluation.DIRs must be (n_vox, 3)")
            # (when a copy was made the caller's array stays untouched, like the reference)
        cfg = evaluation.get_config
        extra = bool(cfg("doSaveModulatedMaps")) if self.id == "NODDI" else bool(cfg("doSaveCorrectedDWI")) if self.id == "FreeWater" else False
        # "amx_exact" is not a reference key: it selects the bit-reproducible kernels (AMX_FLAG_EXACT)
        res = plan.fit(y, dirs, self.solver_params["lambda1"], self.solver_params["lambda2"], rmse=bool(cfg("doComputeRMSE")),
                       nrmse=bool(cfg("doComputeNRMSE")), extra=extra, exact=bool(cfg("amx_exact")))
        return res


class NODDI(BaseModel):
    """NODDI (``amico/models.pyx:655-991``): NNLS -> non-negative elastic net -> NNLS on the support."""

    def __init__(self):
        super().__init__()
        self.id = "NODDI"
        self.name = "NODDI"
        self.maps_name = ["NDI", "ODI", "FWF"]
        self.maps_descr = ["Neurite Density Index", "Orientation Dispersion Index", "Free Water Fraction"]
        self.set()

    def set(self, dPar=1.7E-3, dIso=3.0E-3, IC_VFs=np.linspace(0.1, 0.99, 12),
            IC_ODs=np.hstack((np.array([0.03, 0.06]), np.linspace(0.09, 0.99, 10))), isExvivo=False):
        self.dPar = dPar
        self.dIso = dIso
        self.IC_VFs = np.array(IC_VFs) if isinstance(IC_VFs, list) else IC_VFs
        self.IC_ODs = np.array(IC_ODs) if isinstance(IC_ODs, list) else IC_ODs
        self.isExvivo = isExvivo
        if isExvivo:
            self.maps_name.append("dot")
            self.maps_descr.append("Dot volume fraction")

    def get_params(self):
        return {"id": self.id, "name": self.name, "dPar": self.dPar, "dIso": self.dIso, "IC_VFs": self.IC_VFs,
                "IC_ODs": self.IC_ODs, "isExvivo": self.isExvivo}

    def set_solver(self, lambda1=5e-1, lambda2=1e-3):
        super().set_solver()
        self.solver_params["lambda1"] = lambda1
        self.solver_params["lambda2"] = lambda2

    def _model_params(self):
        return {"isExvivo": bool(self.isExvivo)}

    def _atoms(self):
        for od in self.IC_ODs:  # amico/models.pyx:736-745: kappa outer, v_ic inner, isotropic last
            kappa = 1.0 / np.tan(od * np.pi / 2.0)
            for vf in self.IC_VFs:
                yield _sig.atom_noddi(self.scheme, self.dPar, kappa, vf)
        yield _sig.atom_ball(self.scheme, self.dIso)

    def resample(self, in_path, idx_out, Ylm_out, doMergeB0, ndirs):
        """``amico/models.pyx:754-792``."""
        n_wm = len(self.IC_ODs) * len(self.IC_VFs)
        nS, merge_idx = self._merge_idx(doMergeB0)
        K = {"model": self.id}
        K["wm"] = self._resample_atoms(in_path, 0, n_wm, False, idx_out, Ylm_out, merge_idx, ndirs)
        K["iso"] = self._resample_atoms(in_path, n_wm, 1, True, idx_out, Ylm_out, merge_idx, ndirs)[0]
        K["kappa"] = np.repeat(1.0 / np.tan(np.asarray(self.IC_ODs) * np.pi / 2.0), len(self.IC_VFs)).astype(np.float32)
        K["icvf"] = np.tile(np.asarray(self.IC_VFs), len(self.IC_ODs)).astype(np.float32)
        rows = slice(1, None) if doMergeB0 else self.scheme.dwi_idx
        K["norms"] = np.zeros((self.scheme.dwi_count, n_wm))
        K["norms"][:, :] = 1.0 / np.array([np.linalg.norm(K["wm"][k, 0, rows]) for k in range(n_wm)])  # LUT direction 0 only
        return K


class FreeWater(BaseModel):
    """Free-Water (``amico/models.pyx:994-1286``): one non-negative elastic net on [zeppelins | balls]."""

    def __init__(self):
        super().__init__()
        self.id = "FreeWater"
        self.name = "Free-Water"
        self.set()

    def set(self, d_par=None, d_perps=None, d_isos=None, type="Human"):
        self.type = type
        if self.type == "Mouse":
            self.maps_name = ["FiberVolume", "FW", "FW_blood", "FW_csf"]
            self.maps_descr = ["fiber volume fraction", "Isotropic free-water volume fraction", "FW blood", "FW csf"]
            self.d_par = 1.0E-3 if d_par is None else d_par
            self.d_perps = np.linspace(0.15, 0.55, 10) * 1E-3 if d_perps is None else d_perps
            self.d_isos = [1.5E-3, 3E-3] if d_isos is None else d_isos
        else:
            self.maps_name = ["FiberVolume", "FW"]
            self.maps_descr = ["fiber volume fraction", "Isotropic free-water volume fraction"]
            self.d_par = 1.0E-3 if d_par is None else d_par
            self.d_perps = np.linspace(0.1, 1.0, 10) * 1E-3 if d_perps is None else d_perps
            self.d_isos = [2.5E-3] if d_isos is None else d_isos

    def get_params(self):
        return {"id": self.id, "name": self.name, "d_par": self.d_par, "d_perps": self.d_perps, "d_isos": self.d_isos,
                "type": self.type}

    def set_solver(self, lambda1=0.0, lambda2=1e-3):
        # the reference's 'Mouse' override assigns a local after storing (models.pyx:1084-1085): no effect
        super().set_solver()
        self.solver_params["lambda1"] = lambda1
        self.solver_params["lambda2"] = lambda2

    def _model_params(self):
        return {"type": self.type}

    def _atoms(self):
        for d in self.d_perps:
            yield _sig.atom_zeppelin(self.scheme, self.d_par, d)
        for d in self.d_isos:
            yield _sig.atom_ball(self.scheme, d)

    def resample(self, in_path, idx_out, Ylm_out, doMergeB0, ndirs):
        """``amico/models.pyx:1113-1144``."""
        nS, merge_idx = self._merge_idx(doMergeB0)
        n_perp, n_iso = len(self.d_perps), len(self.d_isos)
        return {"model": self.id,
                "D": self._resample_atoms(in_path, 0, n_perp, False, idx_out, Ylm_out, merge_idx, ndirs),
                "CSF": self._resample_atoms(in_path, n_perp, n_iso, True, idx_out, Ylm_out, merge_idx, ndirs)}


class CylinderZeppelinBall(BaseModel):
    """Cylinder-Zeppelin-Ball / ActiveAx-style (``amico/models.pyx:374-652``)."""

    def __init__(self):
        super().__init__()
        self.id = "CylinderZeppelinBall"
        self.name = "Cylinder-Zeppelin-Ball"
        self.maps_name = ["v", "a", "d"]
        self.maps_descr = ["Intra-cellular volume fraction", "Mean axonal diameter", "Axonal density"]
        self.isExvivo = False  # read but never set by the reference (models.pyx:435, :549): treated as False
        self.set()

    def set(self, d_par=0.6E-3, Rs=np.concatenate(([0.01], np.linspace(0.5, 8.0, 20))) * 1E-6,
            d_perps=np.array([1.19E-3, 0.85E-3, 0.51E-3, 0.17E-3]), d_isos=np.array([2.0E-3])):
        self.d_par = d_par
        self.Rs = np.array(Rs)
        self.d_perps = np.array(d_perps)
        self.d_isos = np.array(d_isos)

    def get_params(self):
        return {"id": self.id, "name": self.name, "d_par": self.d_par, "Rs": self.Rs, "d_perps": self.d_perps,
                "d_isos": self.d_isos, "isExvivo": self.isExvivo}

    def set_solver(self, lambda1=0.0, lambda2=4.0):
        super().set_solver()
        self.solver_params["lambda1"] = lambda1
        self.solver_params["lambda2"] = lambda2

    def _model_params(self):
        return {"Rs": np.asarray(self.Rs, dtype=np.float64)}

    def _atoms(self):
        for R in self.Rs:
            yield _sig.atom_cylinder(self.scheme, self.d_par, R)
        for d in self.d_perps:
            yield _sig.atom_zeppelin(self.scheme, self.d_par, d)
        for d in self.d_isos:
            yield _sig.atom_ball(self.scheme, d)

    def resample(self, in_path, idx_out, Ylm_out, doMergeB0, ndirs):
        """``amico/models.pyx:482-523``."""
        nS, merge_idx = self._merge_idx(doMergeB0)
        n_rs, n_perp, n_iso = len(self.Rs), len(self.d_perps), len(self.d_isos)
        return {"model": self.id,
                "wmr": self._resample_atoms(in_path, 0, n_rs, False, idx_out, Ylm_out, merge_idx, ndirs),
                "wmh": self._resample_atoms(in_path, n_rs, n_perp, False, idx_out, Ylm_out, merge_idx, ndirs),
                "iso": self._resample_atoms(in_path, n_rs + n_perp, n_iso, True, idx_out, Ylm_out, merge_idx, ndirs)}


class SANDI(BaseModel):
    """SANDI (``amico/models.pyx:1343-1627``): one shared, column-normalised dictionary; no direction LUT."""

    def __init__(self):
        super().__init__()
        self.id = "SANDI"
        self.name = "SANDI"
        self.maps_name = ["fsoma", "fneurite", "fextra", "Rsoma", "Din", "De"]
        self.maps_descr = ["Intra-soma volume fraction", "Intra-neurite volume fraction", "Extra-cellular volume fraction",
                           "Apparent soma radius", "Neurite axial diffusivity", "Extra-cellular mean diffusivity"]
        self.set()

    def set(self, d_is=3.0E-3, Rs=np.linspace(1.0, 12.0, 5) * 1E-6, d_in=np.linspace(0.25, 3.0, 5) * 1E-3,
            d_isos=np.linspace(0.25, 3.0, 5) * 1E-3):
        self.d_is = d_is
        self.Rs = np.array(Rs)
        self.d_in = np.array(d_in)
        self.d_isos = np.array(d_isos)

    def get_params(self):
        return {"id": self.id, "name": self.name, "d_is": self.d_is, "Rs": self.Rs, "d_in": self.d_in, "d_isos": self.d_isos}

    def set_solver(self, lambda1=0.0, lambda2=5.0E-3):
        super().set_solver()
        self.solver_params["lambda1"] = lambda1
        self.solver_params["lambda2"] = lambda2

    def _model_params(self):
        return {"Rs": np.asarray(self.Rs, dtype=np.float64), "d_in": np.asarray(self.d_in, dtype=np.float64),
                "d_isos": np.asarray(self.d_isos, dtype=np.float64)}

    def _atoms(self):
        for R in self.Rs:
            yield _sig.atom_sphere(self.scheme, self.d_is, R)
        for d in self.d_in:
            yield _sig.atom_astrosticks(self.scheme, d)
        for d in self.d_isos:
            yield _sig.atom_ball(self.scheme, d)

    def resample(self, in_path, idx_out, Ylm_out, doMergeB0, ndirs):
        """``amico/models.pyx:1446-1486``: every atom is isotropic; columns are L2-normalised."""
        n_atoms = len(self.Rs) + len(self.d_in) + len(self.d_isos)
        nS, merge_idx = self._merge_idx(doMergeB0)
        sig = self._resample_atoms(in_path, 0, n_atoms, True, idx_out, Ylm_out, merge_idx, ndirs)  # (n_atoms, nS) float32
        K = {"model": self.id, "signal": np.zeros((nS, n_atoms), dtype=np.float64, order="F"), "norms": np.zeros(n_atoms)}
        for k in range(n_atoms):
            K["norms"][k] = 1.0 / np.linalg.norm(sig[k])
            K["signal"][:, k] = sig[k] * K["norms"][k]
        return K
