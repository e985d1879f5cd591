"""``Evaluation``: the reference's ``amico.Evaluation`` flow around ``model.fit()`` with the volume resident in HBM.

Mirrors the part of ``amico/core.py`` that sits either side of the per-voxel fit:

* ``load_data``  -- NaN policy, b0 normalisation, b0 merge, directional average (``core.py:151-156, 209-278``),
                    mask compaction and ``y < 0 -> 0`` (``core.py:451-452``): one streaming kernel (``amx_preprocess``);
* ``fit``        -- principal directions (``core.py:430-436, 456-458``; dipy ``TensorModel(OLS)``): ``amx_dti_directions``;
                    ``model.fit`` (``core.py:465``): ``amx_fit`` on device pointers; result scatter (``core.py:472-498``):
                    ``amx_scatter_maps``.

The raw volume is uploaded once; ``y``, ``DIRs`` and the estimates never leave the GPU until ``RESULTS`` is read.  File
I/O (NIfTI, scheme text files), kernel generation/resampling and saving stay with the caller: ``load_data`` takes arrays,
``load_kernels`` takes the ``KERNELS`` dict a ``<Model>.resample`` produced.  torch is used for device memory and streams
only.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import pickle
import time

import numpy as np

from . import _lib as L
from . import models as _models
from . import nifti as _nifti
from .scheme import Scheme

MIN_POSITIVE_SIGNAL = 0.0001  # dipy.reconst.dti.MIN_POSITIVE_SIGNAL
_DIPY_B0_THRESHOLD = 50       # dipy.core.gradients.gradient_table default


def dipy_gradient_table(bvals, bvecs, b0_threshold=_DIPY_B0_THRESHOLD, atol=1e-2):
    """(bvals, bvecs) as dipy's ``gradient_table(bvals, bvecs)`` -> ``GradientTable`` exposes them (restated from the published
    dipy.core.gradients; dipy is absent here, so unpinned): directions whose norm is not within ``atol`` of 1 are zeroed together
    with their b-value (that is how b0 rows with a (0, 0, 0) direction drop out), a diffusion-weighted row (b > 50) with such a
    direction is an error, and the table then holds ``gradients = b * g`` from which ``bvals = |gradients|``,
    ``bvecs = gradients / bvals`` are derived -- so a low-b volume with a unit direction (HCP-style b = 5) KEEPS its gradient."""
    bvals = np.asarray(bvals, dtype=np.float64)
    g = np.array(bvecs, dtype=np.float64)
    g = np.where(np.isnan(g), 0.0, g)
    close = np.abs(np.sqrt((g * g).sum(axis=1)) - 1.0) <= atol
    if not np.all(close[bvals > b0_threshold]):
        raise ValueError("The vectors in bvecs should be unit (The tolerance can be modified as an input parameter)")
    g = np.where(close[:, None], g, 0.0)
    gradients = (bvals * close)[:, None] * g
    b = np.sqrt((gradients * gradients).sum(axis=1))
    return b, gradients / (b + (b == 0))[:, None]


def dti_design_matrix(bvals, bvecs):
    """dipy's ``design_matrix(gtab)`` (lower-triangular order Dxx, Dxy, Dyy, Dxz, Dyz, Dzz, dummy; negated)."""
    bvals, g = dipy_gradient_table(bvals, bvecs)
    B = np.empty((len(bvals), 7))
    B[:, 0] = g[:, 0] * g[:, 0] * bvals
    B[:, 1] = g[:, 0] * g[:, 1] * 2.0 * bvals
    B[:, 2] = g[:, 1] * g[:, 1] * bvals
    B[:, 3] = g[:, 0] * g[:, 2] * 2.0 * bvals
    B[:, 4] = g[:, 1] * g[:, 2] * 2.0 * bvals
    B[:, 5] = g[:, 2] * g[:, 2] * bvals
    B[:, 6] = 1.0
    return -B


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Evaluation:
    """GPU-resident counterpart of ``amico.core.Evaluation`` (``core.py:34-498``) for the load -> fit -> maps flow."""

    def __init__(self, study_path=".", subject=".", output_path=None, device=0):
        """``study_path`` / ``subject`` / ``output_path`` as in ``core.py:36-80``; ``device``: the CUDA device to use."""
        self.device = int(device)
        self.niiDWI = None
        self.scheme = None
        self.model = None
        self.KERNELS = None
        self.htable = None
        self.RESULTS = None
        self.mean_b0s = None
        self.nthreads = None
        self.niiMASK_img = None
        self._y = self._vox_idx = self._dirs = None
        self._dim = None
        self.CONFIG = {"version": "amico_b200", "study_path": study_path, "subject": subject,
                       "DATA_path": os.path.join(study_path, subject), "OUTPUT_path": output_path}
        # defaults of core.py:82-96
        for k, v in (("peaks_filename", None), ("doNormalizeSignal", True), ("doKeepb0Intact", False), ("doComputeRMSE", False),
                     ("doComputeNRMSE", False), ("doSaveModulatedMaps", False), ("doSaveCorrectedDWI", False),
                     ("doMergeB0", False), ("doDebiasSignal", False), ("DWI-SNR", None), ("doDirectionalAverage", False),
                     ("nthreads", -1), ("DTI_fit_method", "OLS"), ("BLAS_nthreads", 1), ("ndirs", 500)):
            self.CONFIG[k] = v

    def set_config(self, key, value):
        self.CONFIG[key] = value

    def get_config(self, key):
        return self.CONFIG.get(key)

    # ------------------------------------------------------------------ load_data (core.py:107-283)
    def load_data(self, dwi, scheme, mask=None, b0_thr=0, b0_min_signal=0, replace_bad_voxels=None, shard=None):
        """See ``_load_data_local``.  With ``shard=(rank, world)`` the ranks first agree that every slab loaded (a rank that raised
        -- NaN policy, geometry mismatch -- would otherwise leave the others waiting in the gathers of ``fit``)."""
        if shard is None or int(shard[1]) <= 1:
            return self._load_data_local(dwi, scheme, mask, b0_thr, b0_min_signal, replace_bad_voxels, shard)
        import torch
        import torch.distributed as dist
        err = None
        try:
            self._load_data_local(dwi, scheme, mask, b0_thr, b0_min_signal, replace_bad_voxels, shard)
        except Exception as e:  # noqa: BLE001 -- re-raised below, on every rank
            err = e
        flag = torch.tensor([0 if err is None else 1], dtype=torch.int32,
                            device=torch.device("cuda", self.device) if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            if err is not None:
                raise err
            raise RuntimeError("loading the data failed on another rank's voxel slab")
        return self

    def _load_data_local(self, dwi, scheme, mask=None, b0_thr=0, b0_min_signal=0, replace_bad_voxels=None, shard=None):
        """``dwi``: file name, (X, Y, Z, nS) array (numpy, or a CUDA float32 torch tensor); ``scheme``: file name, a ``Scheme``
        or an Nx4 / Nx7 table; ``mask``: file name, (X, Y, Z) array or None.  Same pre-processing, same float32 arithmetic as
        the reference.  ``shard=(rank, world)``: this process takes one contiguous slab of the flat voxel list (multi-GPU: one
        process per GPU, every rank sees the same inputs; ``fit`` gathers the volumes on rank 0)."""
        import torch
        lib = L.load()
        if self.get_config("doDebiasSignal"):
            raise NotImplementedError("doDebiasSignal (Rician debiasing, amico/core.py:201-207) is not part of the accelerated path")
        data_path = self.get_config("DATA_path")
        if isinstance(dwi, (str, os.PathLike)):  # file names relative to the subject folder, like core.py:132-141
            fn = os.path.join(data_path, dwi)
            if not os.path.isfile(fn):
                raise FileNotFoundError("DWI file not found")
            self.set_config("dwi_filename", dwi)
            self.niiDWI = _nifti.load(fn)
            dwi = self.niiDWI  # stays in its on-disk dtype and order: converted and transposed on the GPU below
            self.set_config("pixdim", self.niiDWI.zooms[:3])
        if isinstance(scheme, (str, os.PathLike)):
            fn = os.path.join(data_path, scheme)
            if not os.path.isfile(fn):
                raise FileNotFoundError("SCHEME file not found")
            self.set_config("scheme_filename", scheme)
            scheme = _nifti.load_scheme_table(fn)
        if isinstance(mask, (str, os.PathLike)):
            fn = os.path.join(data_path, mask)
            if not os.path.isfile(fn):
                raise FileNotFoundError("MASK file not found")
            self.set_config("mask_filename", mask)
            mask = _nifti.load(fn).get_fdata()
        self.set_config("b0_thr", b0_thr)
        self.set_config("b0_min_signal", b0_min_signal)
        self.set_config("replace_bad_voxels", replace_bad_voxels)
        self.scheme = scheme if isinstance(scheme, Scheme) else Scheme(np.asarray(scheme), b0_thr)
        sch = self.scheme
        dev = torch.device("cuda", self.device)
        # Voxel order inside the pipeline: C order (z fastest) for arrays, the file's own order (x fastest) for NIfTI input --
        # voxels are independent, so only the flattening of the mask and the final reshape of the volumes depend on it.
        self._forder = isinstance(dwi, _nifti.NiftiImage)
        if self._forder:
            img = dwi
            if img.ndim != 4:
                raise ValueError("DWI file is not a 4D image")
            if sch.nS != img.shape[3]:
                raise ValueError("Scheme does not match with DWI data")
            self._dim = tuple(int(d) for d in img.shape[:3])
            n_total = int(np.prod(self._dim))
            code = {v: k for k, v in _nifti._DTYPES.items()}[img.data.dtype.type]
            raw = np.ascontiguousarray(img.data.ravel(order="K")).view(np.uint8)  # the data block as stored: [nS][n_total]
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # read-only buffer: it is only copied to the device
                d_raw = torch.from_numpy(raw).to(dev)
            vol = torch.empty((n_total, sch.nS), dtype=torch.float32, device=dev)
            L.check(lib.amx_volume_to_voxel_major(self.device, L.SPACE_DEVICE, d_raw.data_ptr(), code, n_total, sch.nS, img.scl_slope,
                                                  img.scl_inter, vol.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
            del d_raw
        else:
            if isinstance(dwi, torch.Tensor):
                vol = dwi.to(device=dev, dtype=torch.float32).contiguous()
            else:
                host = np.ascontiguousarray(dwi, dtype=np.float32)  # core.py:136
                vol = torch.from_numpy(host).to(dev, non_blocking=False)
            if vol.ndim != 4:
                raise ValueError("DWI file is not a 4D image")
            if sch.nS != vol.shape[3]:
                raise ValueError("Scheme does not match with DWI data")
            self._dim = tuple(int(s) for s in vol.shape[:3])
            n_total = int(np.prod(self._dim))
        self.set_config("dim", self._dim)
        order = "F" if self._forder else "C"
        self._n_total_full = n_total
        self._shard = None
        if shard is not None and int(shard[1]) > 1:
            from .parallel import shard_bounds
            self._shard = (int(shard[0]), int(shard[1]))
            i0, i1 = shard_bounds(n_total, self._shard[1], self._shard[0])
            self._slab = (i0, i1)
        if mask is not None:
            mask_img = np.asarray(mask).astype(np.uint8)  # core.py:181
            if mask_img.ndim != 3:
                raise ValueError("MASK file is not a 3D image")
            if mask_img.shape != self._dim:
                raise ValueError("MASK geometry does not match with DWI data")
            n_kept = int(np.count_nonzero(mask_img == 1))
            d_mask = torch.from_numpy(np.ascontiguousarray(mask_img.reshape(-1, order=order))).to(dev)
        else:
            mask_img = np.ones(self._dim)  # core.py:190
            n_kept, d_mask = n_total, None
        self.niiMASK_img = mask_img
        flags = 0
        if self.get_config("doNormalizeSignal"):
            flags |= L.PRE_NORMALIZE
        if self.get_config("doMergeB0"):
            flags |= L.PRE_MERGE_B0
        if self.get_config("doDirectionalAverage"):
            flags |= L.PRE_DIR_AVG
        if replace_bad_voxels is not None:
            flags |= L.PRE_REPLACE_BAD
        stream = torch.cuda.current_stream(dev).cuda_stream
        b0_idx, dwi_idx = _i32(sch.b0_idx), _i32(sch.dwi_idx)
        want_b0 = bool(flags & L.PRE_NORMALIZE) and sch.b0_count > 0
        thr = 0.0
        if want_b0 and b0_min_signal != 0:
            # threshold of core.py:216: needs the global mean of the positive b0 means -> one extra streaming pass
            mb = torch.empty(n_total, dtype=torch.float32, device=dev)
            L.check(lib.amx_mean_b0(L.SPACE_DEVICE, self.device, vol.data_ptr(), n_total, sch.nS, b0_idx.ctypes.data, len(b0_idx),
                                    mb.data_ptr(), stream))
            nf = mb.cpu().numpy().reshape(self._dim, order=order)  # numpy's float32 mean depends on the C-order scan
            with np.errstate(all="ignore"):
                thr = float(np.float32(b0_min_signal * nf[nf > 0].mean()))
        if self._shard is not None:  # from here on only this rank's slab of the flat voxel list
            i0, i1 = self._slab
            vol = vol.reshape(-1, sch.nS)[i0:i1].clone()  # own allocation: the kernels want a 16-byte aligned base
            n_total = i1 - i0
            if d_mask is not None:
                d_mask = d_mask[i0:i1].contiguous()
                n_kept = int(np.count_nonzero(mask_img.reshape(-1, order=order)[i0:i1] == 1))
            else:
                n_kept = n_total
        shells = sorted(sch.shells, key=lambda s: s["b"])  # np.argsort(bvals) order (core.py:245)
        sh_idx = _i32(np.concatenate([s["idx"] for s in shells])) if shells else _i32([])
        sh_off = _i32(np.concatenate([[0], np.cumsum([len(s["idx"]) for s in shells])])) if shells else _i32([0])
        m_out = 1 + sch.dwi_count if flags & L.PRE_MERGE_B0 else 1 + len(shells) if flags & L.PRE_DIR_AVG else sch.nS
        y = torch.empty((max(n_kept, 1), m_out), dtype=torch.float32, device=dev)
        vox_idx = torch.empty(max(n_kept, 1), dtype=torch.int32, device=dev)
        mean_b0s = torch.empty(n_total, dtype=torch.float32, device=dev) if want_b0 else None
        a = L.PreArgs()
        a.space, a.device, a.dwi, a.n_total, a.nS = L.SPACE_DEVICE, self.device, vol.data_ptr(), n_total, sch.nS
        a.mask = d_mask.data_ptr() if d_mask is not None else None
        a.b0_idx, a.b0_count, a.dwi_idx, a.dwi_count = b0_idx.ctypes.data, len(b0_idx), dwi_idx.ctypes.data, len(dwi_idx)
        a.shell_idx, a.shell_off, a.n_shells = sh_idx.ctypes.data, sh_off.ctypes.data, len(shells)
        a.flags, a.b0_threshold = flags, thr
        a.replace_bad = float(replace_bad_voxels) if replace_bad_voxels is not None else 0.0
        a.y, a.y_capacity, a.vox_idx = y.data_ptr(), n_kept, vox_idx.data_ptr()
        a.mean_b0s = mean_b0s.data_ptr() if mean_b0s is not None else None
        a.stream = stream
        kept, mo = C.c_int64(0), C.c_int(0)
        if n_total > 0:
            rc = lib.amx_preprocess(C.byref(a), C.byref(kept), C.byref(mo))
            if rc == L.AMX_E_NONFINITE:
                raise FloatingPointError(lib.amx_last_error().decode())
            L.check(rc)
            assert kept.value == n_kept and mo.value == m_out
        # (an empty slab -- more ranks than voxels -- has nothing to pre-process but still takes part in the gathers of fit())
        self._y = y[:n_kept]
        self._vox_idx = vox_idx[:n_kept]
        self._mean_b0s_dev = mean_b0s
        self.mean_b0s = None if (mean_b0s is None or self._shard is not None) else mean_b0s.cpu().numpy().reshape(self._dim, order=order)
        self._dirs = None
        self._fit_scheme = sch
        if flags & L.PRE_DIR_AVG:
            # the scheme the fit sees afterwards: one b0 + one row per shell (core.py:236-262)
            tab = np.zeros((1 + len(shells), 7))
            tab[0] = [1, 0, 0, 0, 0, 0, 0]
            for i, s in enumerate(shells):
                tab[i + 1] = [1, 0, 0, s["G"] or 0, s["Delta"] or 0, s["delta"] or 0, s["TE"] or 0]
            self.scheme = Scheme(tab, b0_thr)
        del vol
        return self

    # ------------------------------------------------------------------ model / kernels
    def set_model(self, model_name):
        """``core.py:286-296``: look the model class up by name."""
        if not hasattr(_models, model_name):
            raise ValueError(f'Model "{model_name}" not recognized')
        self.model = getattr(_models, model_name)()
        self.model.device = self.device
        self.set_config("ATOMS_path", None)  # set by generate_kernels (the reference points it at <study>/kernels/<id> right away)
        self.set_solver()  # default parameters of the fit, like core.py:298
        return self.model

    def set_solver(self, **params):
        """``core.py:300-325``: parameters the model's ``set_solver`` does not know are ignored with a warning, and the accepted
        ones are recorded in ``CONFIG['solver_params']``."""
        if self.model is None:
            raise RuntimeError('Model not set; call "set_model()" method first')
        import inspect
        import warnings
        known = list(inspect.signature(self.model.set_solver).parameters)
        accepted = {}
        for key, value in params.items():
            if key not in known:
                warnings.warn(f"Cannot find the '{key}' solver-parameter for the {self.model.name} model. It will be ignored")
            else:
                accepted[key] = value
        self.model.set_solver(**accepted)
        self.set_config("solver_params", accepted)

    def set_lut(self, directions, htable):
        """The LUT direction set and its 181x181 hash table -- what the reference reads from its package data
        (``amico/directions/ndirs=*.bin``, ``htable_ndirs=*.bin``; ``lut.pyx:50-91``).  Needed by ``generate_kernels`` /
        ``load_kernels()`` without arguments."""
        self._lut_dirs = np.ascontiguousarray(directions, dtype=np.float64)
        self.htable = np.ascontiguousarray(htable, dtype=np.int16)
        self.set_config("ndirs", len(self._lut_dirs))

    def generate_kernels(self, regenerate=False, lmax=12):
        """``core.py:328-365``: write the model's rotated SH-space atoms to ``<study>/kernels/<model id>/A_###.npy`` (skipped when
        they are already there and ``regenerate`` is False)."""
        if self.model is None:
            raise RuntimeError('Model not set; call "set_model()" method first')
        if self.scheme is None:
            raise RuntimeError('Scheme not loaded; call "load_data()" first')
        if getattr(self, "_lut_dirs", None) is None:
            raise RuntimeError('LUT directions not set; call "set_lut()" first')
        from . import lut as _lut
        path = os.path.join(self.get_config("study_path"), "kernels", self.model.id)
        self.set_config("ATOMS_path", path)
        self.set_config("lmax", lmax)
        if os.path.isdir(path) and glob.glob(os.path.join(path, "A_*.npy")) and not regenerate:
            return path
        os.makedirs(path, exist_ok=True)
        for f in glob.glob(os.path.join(path, "*")):
            os.remove(f)
        self.model.scheme = self._fit_scheme if not self.get_config("doDirectionalAverage") else self.scheme
        aux = _lut.precompute_rotation_matrices(lmax, self._lut_dirs)
        idx_in, idx_out = _lut.aux_structures_generate(self.model.scheme, lmax)
        self.model.generate(path, aux, idx_in, idx_out, len(self._lut_dirs))
        return path

    def load_kernels(self, KERNELS=None, htable=None):
        """``core.py:368-404``.  With ``KERNELS``: hand over a dictionary some ``<Model>.resample`` built (and the hash table).
        Without: resample the atoms of ``generate_kernels`` to the subject's scheme on the GPU."""
        if self.model is None:
            raise RuntimeError('Model not set; call "set_model()" method first')
        self.model.scheme = self._fit_scheme if not self.get_config("doDirectionalAverage") else self.scheme
        if KERNELS is None:
            if self.get_config("ATOMS_path") is None:
                raise RuntimeError('Response functions not generated; call "generate_kernels()" first')
            from . import lut as _lut
            idx_out, Ylm_out = _lut.aux_structures_resample(self.model.scheme, self.get_config("lmax") or 12)
            KERNELS = self.model.resample(self.get_config("ATOMS_path"), idx_out, Ylm_out, self.get_config("doMergeB0"),
                                          self.get_config("ndirs"))
        self.KERNELS = KERNELS
        if htable is not None:
            self.htable = htable
        # the plan cache is keyed on the tables' content (models._fingerprint): new tables are noticed at the next fit, equal ones
        # (the usual multi-subject loop re-loading the same kernels) keep their device copies

    # ------------------------------------------------------------------ fit (core.py:407-498)
    @property
    def y(self):
        """``evaluation.y`` of the reference (n_vox, m) float64 -- downloaded on demand."""
        return None if self._y is None else self._rows_in_reference_order(self._y.cpu().numpy().astype(np.float64))

    def _rows_in_reference_order(self, rows):
        """Per-voxel rows in the reference's order (C-order scan of the mask); identity unless the volume came from a file."""
        if not getattr(self, "_forder", False):
            return rows
        f = self._vox_idx.cpu().numpy().astype(np.int64)  # flat index with x fastest
        if getattr(self, "_shard", None) is not None:
            f = f + self._slab[0]  # vox_idx counts from the start of this rank's slab
        X, Y, Z = self._dim
        x, yz = f % X, f // X
        c = (x * Y + yz % Y) * Z + yz // Y
        return rows[np.argsort(c, kind="stable")]

    @property
    def DIRs(self):
        return None if self._dirs is None else self._rows_in_reference_order(self._dirs.cpu().numpy())

    def _dti_tables(self):
        """(design matrix X (m, 7), its pseudo-inverse (7, m)) of the fit scheme, cached per scheme."""
        sch = self._fit_scheme
        if self.get_config("doMergeB0"):  # core.py:432-433
            bvals = np.hstack((0, sch.b[sch.dwi_idx]))
            bvecs = np.vstack((np.zeros((1, 3)), sch.raw[sch.dwi_idx, :3]))
        else:
            bvals, bvecs = sch.b, sch.raw[:, :3]
        key = (bvals.tobytes(), np.ascontiguousarray(bvecs).tobytes())
        if getattr(self, "_dti_key", None) != key:  # the SVD behind pinv costs milliseconds on the host: once per scheme
            X = np.ascontiguousarray(dti_design_matrix(bvals, bvecs), dtype=np.float64)
            self._dti_X, self._dti_W, self._dti_key = X, np.ascontiguousarray(np.linalg.pinv(X), dtype=np.float64), key
        return self._dti_X, self._dti_W

    def estimate_directions(self):
        """Principal directions of every mask voxel (``core.py:456-458``) -> device tensor (n_vox, 3) float64.
        ``DTI_fit_method``: 'OLS' / 'LS' (log-linear least squares, the reference's default) or 'WLS' (dipy's weighted
        variant); 'NLLS', 'RT' / 'RESTORE' (iterative / robust fits, ``core.py:419-429``) are not part of the accelerated path."""
        import torch
        method = self.get_config("DTI_fit_method")
        if method not in ("OLS", "LS", "WLS"):
            raise NotImplementedError(f"DTI fit method {method!r}: only 'OLS' ('LS') and 'WLS' run on the GPU")
        n_vox, m = self._y.shape
        dirs = torch.empty((n_vox, 3), dtype=torch.float64, device=self._y.device)
        X, W = self._dti_tables()
        stream = torch.cuda.current_stream(self._y.device).cuda_stream
        if method == "WLS":
            L.check(L.load().amx_dti_directions_wls(self.device, L.SPACE_DEVICE, self._y.data_ptr(), L.F32, n_vox, m, W.ctypes.data,
                                                    X.ctypes.data, MIN_POSITIVE_SIGNAL, dirs.data_ptr(), stream))
        else:
            W6 = np.ascontiguousarray(W[:6])
            L.check(L.load().amx_dti_directions(self.device, L.SPACE_DEVICE, self._y.data_ptr(), L.F32, n_vox, m, W6.ctypes.data,
                                                MIN_POSITIVE_SIGNAL, dirs.data_ptr(), stream))
        torch.cuda.current_stream(self._y.device).synchronize()  # the tables are host memory read by an async copy
        self._dirs = dirs
        return dirs

    def fit(self):
        import torch
        if self._y is None:
            raise RuntimeError('Data not loaded; call "load_data()" first')
        if self.model is None:
            raise RuntimeError('Model not set; call "set_model()" first')
        if self.KERNELS is None:
            raise RuntimeError('Response functions not generated; call "generate_kernels()" and "load_kernels()" first')
        if self.KERNELS["model"] != self.model.id:
            raise RuntimeError("Response functions were not created with the same model")
        if self.get_config("peaks_filename"):
            # the reference leaves DIRs 4-D on this path (core.py:440-445), which its own chunked fit cannot consume (SURVEY 8a vi)
            raise NotImplementedError('"peaks_filename" is not supported: directions come from the DTI fit')
        if self._shard is None:
            out, res = self._fit_local()
        else:
            # Every rank must reach the gathers below: a rank whose slab fails (NaN policy, LUT range, ...) would otherwise leave the
            # others blocked in the collective.  Agree on the status first; the failing rank re-raises its own exception.
            import torch.distributed as dist
            err = None
            try:
                out, res = self._fit_local()
            except Exception as e:  # noqa: BLE001 -- re-raised below, on every rank
                err, out, res = e, None, None
            flag = torch.tensor([0 if err is None else 1], dtype=torch.int32,
                                device=self._y.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()):
                if err is not None:
                    raise err
                raise RuntimeError("the fit failed on another rank's voxel slab")
        dev = self._y.device
        cfg = self.get_config
        if self._shard is not None:  # one gather per volume: rank 0 ends up with the whole thing, in rank (= voxel) order
            from .parallel import gather_maps
            out = {k: gather_maps(v, self._shard[0], self._shard[1]) for k, v in sorted(out.items())}
            if self._shard[0] != 0:
                self.RESULTS = None
                self._last_fit = res
                return None
        self.RESULTS = {}
        for k, v in out.items():
            a = v.cpu().numpy()
            tail = (a.shape[1],) if k not in ("RMSE", "NRMSE") else ()
            if self._forder:  # memory [z][y][x][c] -> logical (x, y, z, c) as a view
                a = a.reshape(self._dim[::-1] + tail)
                self.RESULTS[k] = a.transpose((2, 1, 0, 3) if tail else (2, 1, 0))
            else:
                self.RESULTS[k] = a.reshape(self._dim + tail)
        self._last_fit = res
        return self.RESULTS

    def _fit_local(self):
        """Directions, fit and scatter of this process's voxels -> ({name: device volume (n_total, k) float32}, fit result)."""
        import torch
        dev = self._y.device
        lib = L.load()
        t = time.time()
        n_vox = self._y.shape[0]
        if not self.get_config("doDirectionalAverage") and self.model.id != "SANDI":
            if n_vox > 0:
                self.estimate_directions()
            else:
                self._dirs = torch.empty((0, 3), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        self.set_config("dirs_precomputing_time", time.time() - t)
        t = time.time()
        if not hasattr(self.model, "solver_params"):
            self.model.set_solver()
        plan = self.model._get_plan(self)
        cfg = self.get_config
        extra = bool(cfg("doSaveModulatedMaps")) if self.model.id == "NODDI" else bool(cfg("doSaveCorrectedDWI")) if self.model.id == "FreeWater" else False
        # the fit flips the directions in place to the y >= 0 hemisphere; RESULTS['DIRs'] holds the DTI output (core.py:477-478)
        fit_dirs = None if self._dirs is None else self._dirs.clone()
        res = plan.fit(self._y, fit_dirs, self.model.solver_params["lambda1"], self.model.solver_params["lambda2"],
                       rmse=bool(cfg("doComputeRMSE")), nrmse=bool(cfg("doComputeNRMSE")), extra=extra, exact=bool(cfg("amx_exact")))
        torch.cuda.synchronize(dev)
        self.set_config("fit_time", time.time() - t)
        # ---- store results (core.py:469-498)
        n_total = int(np.prod(self._dim)) if self._shard is None else self._slab[1] - self._slab[0]
        stream = torch.cuda.current_stream(dev).cuda_stream

        def scatter(values, k):
            vol = torch.zeros((n_total, k), dtype=torch.float32, device=dev) if n_vox == 0 else torch.empty((n_total, k), dtype=torch.float32, device=dev)
            if n_vox > 0:
                L.check(lib.amx_scatter_maps(self.device, L.SPACE_DEVICE, values.data_ptr(), n_vox, k, self._vox_idx.data_ptr(),
                                             vol.data_ptr(), n_total, stream))
            return vol

        out = {"MAPs": scatter(res["estimates"], len(self.model.maps_name))}
        if self._dirs is not None:
            out["DIRs"] = scatter(self._dirs, 3)
        if "rmse" in res:
            out["RMSE"] = scatter(res["rmse"], 1)
        if "nrmse" in res:
            out["NRMSE"] = scatter(res["nrmse"], 1)
        if "estimates_mod" in res:
            out["MAPs_mod"] = scatter(res["estimates_mod"], 2)
        if "y_corrected" in res:
            yc = res["y_corrected"]
            sch = self._fit_scheme
            if cfg("doNormalizeSignal") and sch.b0_count > 0:  # core.py:493-494
                mb = self._mean_b0s_dev[self._vox_idx.long()].to(torch.float64).reshape(-1, 1)
                yc = yc * mb
                if cfg("doKeepb0Intact"):  # core.py:495-496
                    b0 = torch.as_tensor(sch.b0_idx, device=dev)
                    yc[:, b0] = self._y[:, b0].to(torch.float64) * mb
            elif cfg("doKeepb0Intact") and sch.b0_count > 0:
                raise NotImplementedError("doKeepb0Intact without doNormalizeSignal reads mean_b0s the reference never sets")
            out["DWI_corrected"] = scatter(yc.contiguous(), yc.shape[1])
        torch.cuda.synchronize(dev)
        return out, res

    # ------------------------------------------------------------------ save_results (core.py:501-648)
    def save_results(self, path_suffix=None):
        """Write ``config.pickle``, ``fit_dir.nii.gz``, ``fit_<map>.nii.gz`` (+ RMSE / NRMSE / modulated maps / corrected DWI)
        with the reference's file names, header fields and output-folder rule."""
        if self.RESULTS is None:
            raise RuntimeError('Model not fitted to the data; call "fit()" first')
        if self.get_config("OUTPUT_path") is None:
            rel = os.path.join("AMICO", self.model.id) + (("_" + path_suffix) if path_suffix else "")
            self.RESULTS["RESULTS_path"] = rel
            out = os.path.join(self.get_config("DATA_path"), rel)
        else:
            out = self.get_config("OUTPUT_path") + (("_" + path_suffix) if path_suffix else "")
            self.RESULTS["RESULTS_path"] = out
        if not os.path.exists(out):
            os.makedirs(out)
        else:
            for f in glob.glob(os.path.join(out, "*")):
                os.remove(f)
        with open(os.path.join(out, "config.pickle"), "wb+") as fid:
            pickle.dump(self.CONFIG, fid, protocol=2)
        like = self.niiDWI
        ver = self.get_config("version")

        def put(name, img, cal_min, cal_max, descrip=None):
            _nifti.save(os.path.join(out, name), img, like=like, cal_min=cal_min, cal_max=cal_max, descrip=descrip)

        R = self.RESULTS
        if not self.get_config("doDirectionalAverage") and "DIRs" in R:
            put("fit_dir.nii.gz", R["DIRs"], -1, 1)
        if self.get_config("doComputeRMSE"):
            put("fit_RMSE.nii.gz", R["RMSE"], 0, 1)
        if self.get_config("doComputeNRMSE"):
            put("fit_NRMSE.nii.gz", R["NRMSE"], 0, 1)
        if self.get_config("doSaveCorrectedDWI") and self.model.name == "Free-Water":
            put("DWI_corrected.nii.gz", R["DWI_corrected"], 0, 1)
        for i, name in enumerate(self.model.maps_name):
            img = R["MAPs"][:, :, :, i]
            put(f"fit_{name}.nii.gz", img, img.min(), img.max(), f"{self.model.maps_descr[i]} (AMICO v{ver})")
        if self.get_config("doSaveModulatedMaps") and self.model.name == "NODDI":
            for i in range(2):
                img = R["MAPs_mod"][:, :, :, i]
                put(f"fit_{self.model.maps_name[i]}_modulated.nii.gz", img, img.min(), img.max(),
                    f"{self.model.maps_descr[i]} modulated (AMICO v{ver})")
        return out
