"""CPU oracle of what ``amico.Evaluation`` does either side of ``model.fit()``.

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/oracle.py``): imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs, never by the ``amico_b200`` package.

* ``preprocess``      restates ``Evaluation.load_data`` (amico/core.py:151-156, 209-278) and the ``y`` extraction of
                      ``Evaluation.fit`` (amico/core.py:451-452) with the very same numpy statements, so float32 rounding,
                      summation order and the in-place aliasing of the directional average are numpy's own.
* ``dti_directions``  restates what amico/core.py:430-436, 456-458 obtains from dipy (absent here, un-vendored:
                      ``dipy>=1.4.1`` in the reference's pyproject.toml): ``TensorModel(gtab, fit_method='OLS').fit(y)
                      .directions`` = design matrix (dipy/reconst/dti.py ``design_matrix``), ``log(max(y, 1e-4))``
                      (``MIN_POSITIVE_SIGNAL``), ``pinv`` least squares (``ols_fit_tensor``), ``eigh`` and the eigenvector
                      of the largest eigenvalue (``decompose_tensor``).  PARITY UNPINNED: written from dipy's published
                      algorithm, not checked against dipy itself.
* ``scatter_maps``    amico/core.py:472-475.
"""
from __future__ import annotations

import numpy as np

MIN_POSITIVE_SIGNAL = 0.0001  # dipy.reconst.dti.MIN_POSITIVE_SIGNAL
B0_THRESHOLD_DIPY = 50        # dipy.core.gradients.gradient_table default b0_threshold


def preprocess(dwi, scheme, mask=None, *, b0_min_signal=0.0, replace_bad_voxels=None, doNormalizeSignal=True,
               doMergeB0=False, doDirectionalAverage=False):
    """``load_data`` pre-processing + ``y`` extraction.  ``dwi``: (X, Y, Z, nS) array-like.

    Returns dict(y=(n_kept, m) float64 >= 0, vox_idx=flat indices of the kept voxels, mean_b0s=(X, Y, Z) float32 or None,
    img=the pre-processed float32 volume, b0_threshold=the float32 threshold used).
    """
    img = np.array(dwi, dtype=np.float32)  # core.py:136 (.astype(np.float32)); a copy: everything below is in place
    if img.ndim != 4:
        raise ValueError("DWI file is not a 4D image")
    if np.isnan(img).any() or np.isinf(img).any():  # core.py:151-156
        if replace_bad_voxels is not None:
            np.nan_to_num(img, copy=False, nan=replace_bad_voxels, posinf=replace_bad_voxels, neginf=replace_bad_voxels)
        else:
            raise FloatingPointError("Nan or Inf values in the raw signal")
    if mask is None:
        mask_img = np.ones(img.shape[:3])  # core.py:190
    else:
        mask_img = np.asarray(mask).astype(np.uint8)  # core.py:181
    mean_b0s = None
    thr = np.float32(0)
    if doNormalizeSignal:  # core.py:209-222
        if scheme.b0_count <= 0:
            raise ValueError("No b0 volume to normalize signal with")
        mean_b0s = np.mean(img[:, :, :, scheme.b0_idx], axis=3)
        norm_factor = mean_b0s.copy()
        with np.errstate(all="ignore"):
            thr = np.float32(b0_min_signal * norm_factor[norm_factor > 0].mean())  # core.py:216
            idx = norm_factor <= thr
            norm_factor[idx] = 1
            norm_factor = 1 / norm_factor
            norm_factor[idx] = 0
            for i in range(scheme.nS):
                img[:, :, :, i] *= norm_factor
    if doMergeB0:  # core.py:224-227
        mean = np.expand_dims(np.mean(img[:, :, :, scheme.b0_idx], axis=3), axis=3)
        img = np.concatenate((mean, img[:, :, :, scheme.dwi_idx]), axis=3)
    if doDirectionalAverage:  # core.py:231-254
        num_shells = len(scheme.shells)
        dir_avg_img = img[:, :, :, :(num_shells + 1)]  # a VIEW, as in the reference
        id_bval = 0
        dir_avg_img[:, :, :, id_bval] = np.mean(img[:, :, :, scheme.b0_idx], axis=3)
        bvals = [shell["b"] for shell in scheme.shells]
        for shell_idx in np.argsort(bvals):
            shell = scheme.shells[shell_idx]
            id_bval += 1
            dir_avg_img[:, :, :, id_bval] = np.mean(img[:, :, :, shell["idx"]], axis=3)
        img = dir_avg_img.astype(np.float32)
    if np.isnan(img).any() or np.isinf(img).any():  # core.py:273-278
        if replace_bad_voxels is not None:
            np.nan_to_num(img, copy=False, nan=replace_bad_voxels, posinf=replace_bad_voxels, neginf=replace_bad_voxels)
        else:
            raise FloatingPointError("Nan or Inf values in the signal after the pre-processing")
    y = img[mask_img == 1, :].astype(np.double)  # core.py:451
    y[y < 0] = 0  # core.py:452
    vox_idx = np.flatnonzero(mask_img.reshape(-1) == 1)
    return {"y": y, "vox_idx": vox_idx, "mean_b0s": mean_b0s, "img": img, "b0_threshold": thr}


def gradient_table(bvals, bvecs, b0_threshold=B0_THRESHOLD_DIPY, atol=1e-2):
    """``gtab.bvals, gtab.bvecs`` of dipy's ``gradient_table(bvals, bvecs)`` (published dipy.core.gradients, restated -- dipy is
    absent, unpinned): non-unit directions (|norm - 1| > atol) are zeroed with their b-value, which must not happen to a
    diffusion-weighted row; ``gradients = b g``; ``bvals = |gradients|``; ``bvecs = gradients / bvals``."""
    bvals = np.asarray(bvals, dtype=np.float64)
    bvecs = np.array(bvecs, dtype=np.float64)
    bvecs = np.where(np.isnan(bvecs), 0, bvecs)
    close = np.abs(np.linalg.norm(bvecs, axis=1) - 1) <= atol
    if not np.all(close[bvals > b0_threshold]):
        raise ValueError("The vectors in bvecs should be unit (The tolerance can be modified as an input parameter)")
    bvecs = np.where(close[:, None], bvecs, 0)
    gradients = (bvals * close)[:, None] * bvecs
    b = np.linalg.norm(gradients, axis=1)
    return b, gradients / (b + (b == 0))[:, None]


def dti_design_matrix(bvals, bvecs):
    """dipy ``design_matrix(gtab)``: columns Dxx, Dxy, Dyy, Dxz, Dyz, Dzz, dummy (lower-triangular order), negated."""
    bvals, bvecs = gradient_table(bvals, bvecs)
    B = np.zeros((len(bvals), 7))
    B[:, 0] = bvecs[:, 0] * bvecs[:, 0] * 1.0 * bvals
    B[:, 1] = bvecs[:, 0] * bvecs[:, 1] * 2.0 * bvals
    B[:, 2] = bvecs[:, 1] * bvecs[:, 1] * 1.0 * bvals
    B[:, 3] = bvecs[:, 0] * bvecs[:, 2] * 2.0 * bvals
    B[:, 4] = bvecs[:, 1] * bvecs[:, 2] * 2.0 * bvals
    B[:, 5] = bvecs[:, 2] * bvecs[:, 2] * 1.0 * bvals
    B[:, 6] = np.ones(len(bvals))
    return -B


def scheme_gradients(scheme, doMergeB0=False):
    """(bvals, bvecs) exactly as amico/core.py:432-435 hands them to ``gradient_table``."""
    if doMergeB0:
        return (np.hstack((0, scheme.b[scheme.dwi_idx])), np.vstack((np.zeros((1, 3)), scheme.raw[scheme.dwi_idx, :3])))
    return scheme.b, scheme.raw[:, :3]


def dti_directions(y, scheme, doMergeB0=False, method="OLS"):
    """Principal eigenvector of the tensor fit per voxel: (n_vox, 3) float64 (sign arbitrary).  ``method``: 'OLS' (dipy
    ``ols_fit_tensor``: pinv(X) log s) or 'WLS' (dipy ``wls_fit_tensor``: w = exp(U U^T log s) with U from the thin SVD of X, then
    pinv(X * w[:, None]) (w * log s)), both restated from the published dipy.reconst.dti (dipy is absent: unpinned)."""
    bvals, bvecs = scheme_gradients(scheme, doMergeB0)
    X = dti_design_matrix(bvals, bvecs)
    data = np.maximum(np.asarray(y, dtype=np.float64), MIN_POSITIVE_SIGNAL)
    log_s = np.log(data)
    if method == "OLS":
        D = np.einsum("ij,nj->ni", np.linalg.pinv(X), log_s)
    elif method == "WLS":
        U = np.linalg.svd(X, full_matrices=False)[0]
        w = np.exp(np.einsum("ij,nj->ni", U @ U.T, log_s))
        D = np.empty((len(log_s), 7))
        for i in range(len(log_s)):
            D[i] = np.linalg.pinv(X * w[i][:, None]) @ (w[i] * log_s[i])
    else:
        raise ValueError(method)
    T = np.empty((len(D), 3, 3))
    T[:, 0, 0], T[:, 0, 1], T[:, 1, 1], T[:, 0, 2], T[:, 1, 2], T[:, 2, 2] = (D[:, k] for k in range(6))
    T[:, 1, 0], T[:, 2, 0], T[:, 2, 1] = T[:, 0, 1], T[:, 0, 2], T[:, 1, 2]
    _, evecs = np.linalg.eigh(T)  # ascending eigenvalues: the principal direction is the last column
    return np.ascontiguousarray(evecs[:, :, 2])


def scatter_maps(values, vox_idx, n_total):
    """``RESULTS['MAPs'][mask==1, :] = estimates`` into a zero float32 volume (core.py:474-475); flat (n_total, k)."""
    values = np.asarray(values)
    vol = np.zeros((n_total, values.shape[1]), dtype=np.float32)
    vol[np.asarray(vox_idx)] = values
    return vol


# --------------------------------------------------------------------------- kernel resampling (row f-3)
def resample_kernel(KRlm, nS, idx_out, Ylm_out, is_isotropic, ndirs):
    """``amico/lut.pyx:274-311`` statement for statement (numpy float32 ``np.dot``)."""
    if not is_isotropic:
        KR = np.ones((ndirs, nS), dtype=np.float32)
        for i in range(ndirs):
            KR[i, idx_out] = np.dot(Ylm_out, KRlm[i, :]).astype(np.float32)
    else:
        KR = np.ones(nS, dtype=np.float32)
        KR[idx_out] = np.dot(Ylm_out, KRlm)
    return KR


def resample_kernel_exact(KRlm, nS, idx_out, Ylm_out, is_isotropic, ndirs):
    """The same projection with the dot products in float64, rounded once: the value every float32 BLAS summation order
    approximates (the GPU kernel must match THIS to 1 ulp)."""
    Y = np.asarray(Ylm_out, dtype=np.float64)
    K = np.asarray(KRlm, dtype=np.float64)
    if not is_isotropic:
        KR = np.ones((ndirs, nS), dtype=np.float32)
        KR[:, idx_out] = (K @ Y.T).astype(np.float32)
    else:
        KR = np.ones(nS, dtype=np.float32)
        KR[idx_out] = (Y @ K).astype(np.float32)
    return KR


def model_resample(model, params, scheme, in_path, idx_out, Ylm_out, doMergeB0, ndirs, kernel=resample_kernel):
    """``<Model>.resample`` of amico/models.pyx (:754-792 NODDI, :1113-1144 FreeWater, :482-523 CylinderZeppelinBall,
    :1446-1486 SANDI) on the ``A_###.npy`` files of ``in_path``."""
    import os
    if doMergeB0:
        nS = 1 + scheme.dwi_count
        merge_idx = np.hstack((scheme.b0_idx[0], scheme.dwi_idx))
    else:
        nS = scheme.nS
        merge_idx = np.arange(nS)
    load = lambda i: np.load(os.path.join(in_path, f"A_{i + 1:03d}.npy"))
    K = {"model": model}
    if model == "NODDI":
        ods, vfs = params["IC_ODs"], params["IC_VFs"]
        n = len(ods) * len(vfs)
        K["wm"] = np.zeros((n, ndirs, nS), dtype=np.float32)
        K["kappa"] = np.zeros(n, dtype=np.float32)
        K["icvf"] = np.zeros(n, dtype=np.float32)
        K["norms"] = np.zeros((scheme.dwi_count, n))
        idx = 0
        for i in range(len(ods)):
            for j in range(len(vfs)):
                K["wm"][idx] = kernel(load(idx), scheme.nS, idx_out, Ylm_out, False, ndirs)[:, merge_idx]
                K["kappa"][idx] = 1.0 / np.tan(ods[i] * np.pi / 2.0)
                K["icvf"][idx] = vfs[j]
                if doMergeB0:
                    K["norms"][:, idx] = 1 / np.linalg.norm(K["wm"][idx, 0, 1:])
                else:
                    K["norms"][:, idx] = 1 / np.linalg.norm(K["wm"][idx, 0, scheme.dwi_idx])
                idx += 1
        K["iso"] = kernel(load(n), scheme.nS, idx_out, Ylm_out, True, ndirs)[merge_idx]
    elif model == "FreeWater":
        n_perp, n_iso = len(params["d_perps"]), len(params["d_isos"])
        K["D"] = np.zeros((n_perp, ndirs, nS), dtype=np.float32)
        K["CSF"] = np.zeros((n_iso, nS), dtype=np.float32)
        for i in range(n_perp):
            K["D"][i] = kernel(load(i), scheme.nS, idx_out, Ylm_out, False, ndirs)[:, merge_idx]
        for i in range(n_iso):
            K["CSF"][i] = kernel(load(n_perp + i), scheme.nS, idx_out, Ylm_out, True, ndirs)[merge_idx]
    elif model == "CylinderZeppelinBall":
        n_rs, n_perp, n_iso = len(params["Rs"]), len(params["d_perps"]), len(params["d_isos"])
        K["wmr"] = np.zeros((n_rs, ndirs, nS), dtype=np.float32)
        K["wmh"] = np.zeros((n_perp, ndirs, nS), dtype=np.float32)
        K["iso"] = np.zeros((n_iso, nS), dtype=np.float32)
        for i in range(n_rs):
            K["wmr"][i] = kernel(load(i), scheme.nS, idx_out, Ylm_out, False, ndirs)[:, merge_idx]
        for i in range(n_perp):
            K["wmh"][i] = kernel(load(n_rs + i), scheme.nS, idx_out, Ylm_out, False, ndirs)[:, merge_idx]
        for i in range(n_iso):
            K["iso"][i] = kernel(load(n_rs + n_perp + i), scheme.nS, idx_out, Ylm_out, True, ndirs)[merge_idx]
    elif model == "SANDI":
        n = len(params["Rs"]) + len(params["d_in"]) + len(params["d_isos"])
        K["signal"] = np.zeros((nS, n), dtype=np.float64, order="F")
        K["norms"] = np.zeros(n, dtype=np.float64)
        for idx in range(n):
            signal = kernel(load(idx), scheme.nS, idx_out, Ylm_out, True, ndirs)[merge_idx].T
            K["norms"][idx] = 1.0 / np.linalg.norm(signal)
            K["signal"][:, idx] = signal * K["norms"][idx]
    else:
        raise ValueError(model)
    return K
