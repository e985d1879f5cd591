"""Kernel resampling: the part of ``amico/lut.pyx`` that turns the rotated SH-space atoms ``generate`` wrote to disk
into the subject's ``KERNELS``, with the projection on the GPU.

* ``real_sh_descoteaux``      -- the real, even-order SH basis dipy's function of that name returns with its legacy
                                 convention (dipy is absent here; restated from its definition: for m < 0
                                 ``sqrt(2) Re Y_l^|m|``, for m = 0 ``Y_l^0``, for m > 0 ``sqrt(2) Im Y_l^m``; orders
                                 l = 0, 2, ..., lmax, m = -l..l -- the ordering ``amico/lut.pyx:129-138`` relies on);
* ``aux_structures_resample`` -- ``amico/lut.pyx:186-224``: ``idx_OUT`` and the block matrix ``Ylm_OUT`` (float32);
* ``resample_kernels``        -- ``amico/lut.pyx:274-311`` for a stack of atoms: one ``amx_resample_kernels`` call.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L


def cart2sphere(x, y, z):
    """dipy.core.geometry.cart2sphere: r, theta (polar angle from +z), phi (azimuth)."""
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    r = np.sqrt(x * x + y * y + z * z)
    with np.errstate(invalid="ignore", divide="ignore"):
        theta = np.arccos(np.divide(z, r, out=np.zeros_like(r), where=r > 0))
    theta = np.where(r > 0, theta, 0.0)
    phi = np.arctan2(y, x)
    return r, theta, phi


def sh_index_list(lmax):
    """(m, l) of every basis function, even l only: l = 0, 2, ..., lmax and m = -l..l (dipy ``sph_harm_ind_list``)."""
    ms, ls = [], []
    for l in range(0, lmax + 1, 2):
        for m in range(-l, l + 1):
            ms.append(m)
            ls.append(l)
    return np.array(ms), np.array(ls)


def real_sh_descoteaux(lmax, theta, phi):
    """Real SH basis matrix (n_points, n_sh) at polar angles ``theta`` and azimuths ``phi``; also returns (m, l)."""
    from scipy.special import sph_harm_y
    m, l = sh_index_list(lmax)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1, 1)
    phi = np.asarray(phi, dtype=np.float64).reshape(-1, 1)
    sh = sph_harm_y(l[None, :], np.abs(m)[None, :], theta, phi)
    real = np.where(m[None, :] > 0, sh.imag, sh.real)
    real = real * np.where(m[None, :] == 0, 1.0, np.sqrt(2.0))
    return real, m, l


def aux_structures_resample(scheme, lmax=12):
    """``amico/lut.pyx:186-224``: (idx_OUT int32 [dwi_count], Ylm_OUT float32 [dwi_count, nSH * n_shells])."""
    n_sh = (lmax + 1) * (lmax + 2) // 2
    idx_out = np.zeros(scheme.dwi_count, dtype=np.int32)
    ylm_out = np.zeros((scheme.dwi_count, n_sh * len(scheme.shells)), dtype=np.float32)
    idx = 0
    for s, shell in enumerate(scheme.shells):
        n = len(shell["idx"])
        idx_out[idx:idx + n] = shell["idx"]
        g = shell["grad"]
        _, theta, phi = cart2sphere(g[:, 0], g[:, 1], g[:, 2])
        ylm_out[idx:idx + n, n_sh * s:n_sh * (s + 1)] = real_sh_descoteaux(lmax, theta, phi)[0]
        idx += n
    return idx_out, ylm_out


def resample_kernels(KRlm, nS, idx_out, Ylm_out, merge_idx=None, device=0):
    """``resample_kernel(...)[..., merge_idx]`` for a stack of atoms.

    ``KRlm``: float32 (n_atoms, ndirs, n_coef) for rotated atoms or (n_atoms, n_coef) for isotropic ones.  Returns the
    float32 array (n_atoms, ndirs, len(merge_idx)) or (n_atoms, len(merge_idx)).  Runs on the GPU (no CPU fallback).
    """
    lib = L.load()
    K = np.ascontiguousarray(KRlm, dtype=np.float32)
    Y = np.ascontiguousarray(Ylm_out, dtype=np.float32)
    io = np.ascontiguousarray(idx_out, dtype=np.int32)
    mi = np.ascontiguousarray(np.arange(nS) if merge_idx is None else merge_idx, dtype=np.int32)
    n_coef = K.shape[-1]
    if Y.shape != (len(io), n_coef):
        # the reference reports a shape mismatch of its np.dot as an outdated LUT (lut.pyx:302-303)
        raise RuntimeError('Outdated LUT. Call "generate_kernels( regenerate=True )" to update the LUT')
    n_rows = int(np.prod(K.shape[:-1]))
    out = np.empty(K.shape[:-1] + (len(mi),), dtype=np.float32)
    L.check(lib.amx_resample_kernels(int(device), L.SPACE_HOST, K.ctypes.data, n_rows, n_coef, Y.ctypes.data, io.ctypes.data,
                                     len(io), mi.ctypes.data, len(mi), int(nS), out.ctypes.data, None))
    return out


# --------------------------------------------------------------------------- kernel generation: rotation in SH space
def precompute_rotation_matrices(lmax, directions):
    """``AUX`` of ``amico/lut.pyx:94-141`` for the LUT direction set ``directions`` (ndirs, 3): ``Ylm_rot`` (the SH basis at every
    LUT direction), ``const`` = sqrt(4 pi / (2l+1)) and ``idx_m0`` (position of the m = 0 function of each order) per basis
    function.  (The reference also stores ``fit``, the least-squares SH fit on its 500-direction grid; the zonal atoms of
    ``amico_b200.signals`` come with their Legendre coefficients, so no fit is needed -- see ``rotate_kernel``.)"""
    directions = np.asarray(directions, dtype=np.float64)
    _, theta, phi = cart2sphere(directions[:, 0], directions[:, 1], directions[:, 2])
    Y, m, l = real_sh_descoteaux(lmax, theta, phi)
    return {"lmax": lmax, "ndirs": len(directions), "Ylm_rot": Y, "const": np.sqrt(4.0 * np.pi / (2.0 * l + 1.0)),
            "idx_m0": ((l * l + l + 2) // 2 - 1).astype(np.int32), "l": l}


def aux_structures_generate(scheme, lmax=12):
    """``amico/lut.pyx:162-183``: sample indices of each shell in the 500-direction high-resolution scheme and SH indices."""
    n_sh = (lmax + 1) * (lmax + 2) // 2
    idx_in = [range(500 * s, 500 * (s + 1)) for s in range(len(scheme.shells))]
    idx_out = [range(n_sh * s, n_sh * (s + 1)) for s in range(len(scheme.shells))]
    return idx_in, idx_out


def rotate_kernel(atom, aux, idx_out, ndirs):
    """``rotate_kernel`` of ``amico/lut.pyx:227-271`` for a zonal atom given by its per-shell Legendre coefficients a_l.

    The reference projects the signal on the SH basis (``K_lm``), keeps the m = 0 coefficient of every order and writes
    ``KRlm[i, (l, m)] = sqrt(4 pi/(2l+1)) K_l0 Y_lm(dir_i)``.  For f(t) = sum_l a_l P_l(t) the m = 0 coefficient is
    ``K_l0 = a_l sqrt(4 pi/(2l+1))`` exactly, so ``KRlm[i, (l, m)] = a_l 4 pi/(2l+1) Y_lm(dir_i)``.  Isotropic atoms: the SH
    coefficients themselves (only l = 0: value * sqrt(4 pi)).  Returns float32 (ndirs, n_sh * n_shells) or (n_sh * n_shells,).
    """
    n_sh = aux["Ylm_rot"].shape[1]
    n = n_sh * len(atom.per_shell)
    if atom.isotropic:
        out = np.zeros(n, dtype=np.float32)
        for s, v in enumerate(atom.per_shell):
            out[idx_out[s][0]] = np.float32(v * np.sqrt(4.0 * np.pi))
        return out
    if aux["ndirs"] != ndirs:
        raise ValueError("AUX was computed for another number of directions")
    order = (aux["l"] // 2).astype(int)  # index of the order l of each basis function in a_l (l = 0, 2, ...)
    out = np.zeros((ndirs, n), dtype=np.float32)
    for s, f in enumerate(atom.per_shell):
        k_l0 = f.coeffs * np.sqrt(4.0 * np.pi / (2.0 * np.arange(0, aux["lmax"] + 1, 2) + 1.0))
        out[:, list(idx_out[s])] = (aux["const"] * k_l0[order])[None, :] * aux["Ylm_rot"]
    return out
