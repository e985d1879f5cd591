"""Synthetic protocols, dictionaries and volumes for the per-voxel fit.

Host-side tooling (numpy/scipy) used by the tests and by ``bench.py``: nothing
here is on the product path.  It builds inputs in exactly the layouts the
reference hands to ``model.fit(evaluation)``:

* ``Scheme``         -- the attributes of ``amico/scheme.py:21-154`` the path reads
                        (``b``, ``b0_idx``, ``dwi_idx``, ``dwi_count``, ``nS``, ``shells``);
* ``lut_directions`` / ``build_htable`` -- a half-sphere direction set and the
                        181x181 one-degree hash table with the semantics of
                        ``amico/directions/htable_ndirs=*.bin`` (nearest LUT
                        direction by |dot|; checked against the reference's own
                        500-direction table in ``tests/test_oracle_cpu.py``);
* ``make_kernels``   -- the ``KERNELS`` dict of each model's ``resample``
                        (``amico/models.pyx:754-792, 1113-1144, 482-523, 1446-1486``);
* ``make_voxels``    -- ``y`` (float32-valued, >= 0) and unit ``DIRs``.

The compartment signals are the standard closed forms (stick / zeppelin /
ball, Watson-dispersed NODDI sticks with tortuosity, Van Gelderen cylinder,
Murday-Cotts sphere, astrosticks).  Every anisotropic atom is a zonal function
of ``t = g . d`` per shell; as in the reference it is band-limited to even
Legendre orders <= 12 before being sampled on the subject's gradients (the
reference does this through an lmax=12 SH fit, ``amico/lut.pyx:227-311``; the
m=0 coefficients it keeps are exactly a Legendre series).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


CONFIGS = {
    # id: (model, volume dims, description)
    1: ("FreeWater", (8, 8, 8)),
    2: ("NODDI", (128, 128, 64)),
    3: ("NODDI", (256, 256, 160)),
    4: ("SANDI", (128, 128, 128)),
    5: ("CylinderZeppelinBall", (200, 200, 200)),
}


# --------------------------------------------------------------------------- scheme
from .scheme import GAMMA, Scheme  # noqa: E402,F401  (re-exported: tests and tools import it from here)


def fibonacci_sphere(n, phase=0.0):
    """n quasi-uniform unit vectors on the full sphere."""
    k = np.arange(n) + 0.5
    z = 1.0 - 2.0 * k / n
    phi = (math.pi * (3.0 - math.sqrt(5.0))) * k + phase
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def make_scheme(cfg):
    """Synthetic protocol of SURVEY 8(d) / BASELINE.md section 3 for config ``cfg``."""
    def shell_dirs(n, s):
        return fibonacci_sphere(n, phase=0.37 * (s + 1))

    if cfg == 1:
        parts = [(1, 0.0), (32, 1000.0)]
    elif cfg == 2:
        parts = [(10, 0.0), (45, 1000.0), (45, 2000.0)]
    elif cfg == 3:
        parts = [(18, 0.0), (135, 1000.0), (135, 2000.0)]
    elif cfg == 4:
        parts = [(12, 0.0), (60, 1000.0), (60, 3000.0), (60, 10000.0)]
    elif cfg == 5:
        parts = [(12, 0.0), (72, 1000.0), (72, 2000.0), (72, 3000.0), (72, 4000.0)]
    else:
        raise ValueError(cfg)
    stejskal = cfg in (4, 5)
    Delta, delta, TE = 0.040, 0.020, 0.080
    rows = []
    s = 0
    for n, b in parts:
        if b == 0.0:
            g = np.zeros((n, 3))
        else:
            g = shell_dirs(n, s)
            s += 1
        if stejskal:
            G = math.sqrt(b * 1e6 / (Delta - delta / 3.0)) / (GAMMA * delta) if b > 0 else 0.0
            rows.append(np.hstack([g, np.tile([G, Delta, delta, TE], (n, 1))]))
        else:
            rows.append(np.hstack([g, np.full((n, 1), b)]))
    return Scheme(np.vstack(rows))


def directional_average_scheme(scheme):
    """Scheme after ``doDirectionalAverage`` (``amico/core.py:232-268``): one b0 + one row per shell."""
    n = 1 + len(scheme.shells)
    raw = np.zeros((n, scheme.raw.shape[1]))
    for i, sh in enumerate(scheme.shells):
        raw[i + 1, 0:3] = [1.0, 0.0, 0.0]
        raw[i + 1, 3:] = scheme.raw[sh["idx"][0], 3:]
    return Scheme(raw)


# --------------------------------------------------------------------------- LUT directions
def lut_directions(ndirs=500):
    """Half-sphere (y >= 0) direction set standing in for ``directions/ndirs=*.bin``."""
    d = fibonacci_sphere(2 * ndirs)
    d = d[d[:, 1] >= 0]
    if len(d) < ndirs:  # pragma: no cover
        raise RuntimeError("direction set too small")
    return np.ascontiguousarray(d[:ndirs])


def build_htable(dirs):
    """int16[181*181]: nearest direction (max |dot|) for each whole-degree (theta, phi)."""
    ang = np.deg2rad(np.arange(181.0))
    T, P = np.meshgrid(ang, ang, indexing="ij")
    v = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3)
    return np.abs(v @ dirs.T).argmax(1).astype(np.int16)


# --------------------------------------------------------------------------- zonal atoms
from .signals import (LMAX, _Atom, _sample, atom_astrosticks, atom_ball, atom_cylinder, atom_noddi, atom_sphere,  # noqa: E402,F401
                      atom_stick, atom_zeppelin)


# --------------------------------------------------------------------------- model parameter grids
def default_params(model):
    """Default physical grids of each reference model (``amico/models.pyx:400-424, 675-706, 1004-1058, 1367-1391``)."""
    if model == "NODDI":
        return dict(dPar=1.7e-3, dIso=3.0e-3, IC_VFs=np.linspace(0.1, 0.99, 12),
                    IC_ODs=np.hstack((np.array([0.03, 0.06]), np.linspace(0.09, 0.99, 10))), isExvivo=False)
    if model == "FreeWater":
        return dict(d_par=1.0e-3, d_perps=np.linspace(0.1, 1.0, 10) * 1e-3, d_isos=[2.5e-3], type="Human")
    if model == "FreeWaterMouse":
        return dict(d_par=1.0e-3, d_perps=np.linspace(0.15, 0.55, 10) * 1e-3, d_isos=[1.5e-3, 3e-3], type="Mouse")
    if model == "CylinderZeppelinBall":
        return dict(d_par=0.6e-3, Rs=np.concatenate(([0.01], np.linspace(0.5, 8.0, 20))) * 1e-6,
                    d_perps=np.array([1.19e-3, 0.85e-3, 0.51e-3, 0.17e-3]), d_isos=np.array([2.0e-3]))
    if model == "SANDI":
        return dict(d_is=3.0e-3, Rs=np.linspace(1.0, 12.0, 5) * 1e-6, d_in=np.linspace(0.25, 3.0, 5) * 1e-3,
                    d_isos=np.linspace(0.25, 3.0, 5) * 1e-3)
    raise ValueError(model)


def make_kernels(model, scheme, dirs, params=None):
    """KERNELS dict in the layout ``<Model>.resample`` returns (no b0 merging)."""
    p = dict(default_params(model))
    if params:
        p.update(params)
    nS, nd = scheme.nS, len(dirs)
    K = {}
    if model == "NODDI":
        K["model"] = "NODDI"
        n_wm = len(p["IC_ODs"]) * len(p["IC_VFs"])
        K["wm"] = np.zeros((n_wm, nd, nS), dtype=np.float32)
        K["kappa"] = np.zeros(n_wm, dtype=np.float32)
        K["icvf"] = np.zeros(n_wm, dtype=np.float32)
        K["norms"] = np.zeros((scheme.dwi_count, n_wm))
        idx = 0
        for od in p["IC_ODs"]:
            kappa = 1.0 / math.tan(od * math.pi / 2.0)
            for vf in p["IC_VFs"]:
                K["wm"][idx] = _sample(atom_noddi(scheme, p["dPar"], kappa, vf), scheme, dirs)
                K["kappa"][idx] = kappa
                K["icvf"][idx] = vf
                K["norms"][:, idx] = 1.0 / np.linalg.norm(K["wm"][idx, 0, scheme.dwi_idx])
                idx += 1
        K["iso"] = _sample(atom_ball(scheme, p["dIso"]), scheme, dirs)
    elif model in ("FreeWater", "FreeWaterMouse"):
        K["model"] = "FreeWater"
        K["D"] = np.stack([_sample(atom_zeppelin(scheme, p["d_par"], d), scheme, dirs) for d in p["d_perps"]])
        K["CSF"] = np.stack([_sample(atom_ball(scheme, d), scheme, dirs) for d in p["d_isos"]])
    elif model == "CylinderZeppelinBall":
        K["model"] = "CylinderZeppelinBall"
        K["wmr"] = np.stack([_sample(atom_cylinder(scheme, p["d_par"], R), scheme, dirs) for R in p["Rs"]])
        K["wmh"] = np.stack([_sample(atom_zeppelin(scheme, p["d_par"], d), scheme, dirs) for d in p["d_perps"]])
        K["iso"] = np.stack([_sample(atom_ball(scheme, d), scheme, dirs) for d in p["d_isos"]])
    elif model == "SANDI":
        K["model"] = "SANDI"
        atoms = [atom_sphere(scheme, p["d_is"], R) for R in p["Rs"]] \
            + [atom_astrosticks(scheme, d) for d in p["d_in"]] + [atom_ball(scheme, d) for d in p["d_isos"]]
        K["signal"] = np.zeros((nS, len(atoms)), dtype=np.float64, order="F")
        K["norms"] = np.zeros(len(atoms))
        for i, a in enumerate(atoms):
            s = _sample(a, scheme, dirs).astype(np.float64)
            K["norms"][i] = 1.0 / np.linalg.norm(s)
            K["signal"][:, i] = s * K["norms"][i]
    else:
        raise ValueError(model)
    return K, p


def dictionary_for_direction(model, K, k):
    """float64 (m, n) dictionary the reference assembles for LUT index k (``models.pyx:905-908`` etc.)."""
    if model == "NODDI":
        return np.hstack([K["wm"][:, k, :].T.astype(np.float64), K["iso"].astype(np.float64)[:, None]])
    if model in ("FreeWater", "FreeWaterMouse"):
        return np.hstack([K["D"][:, k, :].T.astype(np.float64), K["CSF"].T.astype(np.float64)])
    if model == "CylinderZeppelinBall":
        return np.hstack([K["wmr"][:, k, :].T.astype(np.float64), K["wmh"][:, k, :].T.astype(np.float64),
                          K["iso"].T.astype(np.float64)])
    if model == "SANDI":
        return np.asarray(K["signal"], dtype=np.float64)
    raise ValueError(model)


# --------------------------------------------------------------------------- voxels
def lut_index_numpy(dirs_in, htable):
    """Vectorised restatement of ``amico/lut.pyx:314-356`` (does not flip its input)."""
    d = np.array(dirs_in, dtype=np.float64)
    neg = d[:, 1] < 0
    d[neg] *= -1.0
    two_pi = 2.0 * math.pi
    i2 = np.fmod(np.arctan2(d[:, 1], d[:, 0]), two_pi)
    lo = i2 < 0
    i2[lo] = np.fmod(i2[lo] + two_pi, two_pi)
    big = i2 > math.pi
    rxy = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)
    i1 = np.arctan2(rxy, d[:, 2])
    i2[big] = np.fmod(np.arctan2(-d[big, 1], -d[big, 0]), two_pi)
    i1[big] = np.arctan2(rxy[big], -d[big, 2])
    c_round = lambda x: np.sign(x) * np.floor(np.abs(x) + 0.5)
    ii1 = c_round(i1 / math.pi * 180.0).astype(np.int64)
    ii2 = c_round(i2 / math.pi * 180.0).astype(np.int64)
    if ((ii1 < 0) | (ii1 > 180) | (ii2 < 0) | (ii2 > 180)).any():
        raise RuntimeError('"amico.lut.dir_to_lut_idx" index out of bounds')
    return htable[ii1 * 181 + ii2].astype(np.int64)


def make_voxels(model, K, htable, n_vox, seed, snr=30.0, iso_max=0.5):
    """Seeded voxels: ground truth = 1-2 atoms + isotropic fraction, Rician noise at ``snr``.

    Returns ``y`` (n_vox, m) float32 >= 0 and ``dirs`` (n_vox, 3) float64 unit vectors.
    """
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n_vox, 3))
    dirs = v / np.linalg.norm(v, axis=1, keepdims=True)
    f_iso = rng.uniform(0.0, iso_max, n_vox)
    two = rng.random(n_vox) < 0.5
    w1 = np.where(two, rng.uniform(0.3, 0.7, n_vox), 1.0)
    if model == "SANDI":
        A = np.asarray(K["signal"]) / K["norms"][None, :]
        n = A.shape[1]
        j1, j2 = rng.integers(0, n, n_vox), rng.integers(0, n, n_vox)
        sig = w1[:, None] * A[:, j1].T + (1 - w1)[:, None] * A[:, j2].T
    else:
        k = lut_index_numpy(dirs, htable)
        if model == "NODDI":
            rot, iso = K["wm"], K["iso"][None, :]
        elif model in ("FreeWater", "FreeWaterMouse"):
            rot, iso = K["D"], K["CSF"]
        else:
            rot, iso = np.concatenate([K["wmr"], K["wmh"]]), K["iso"]
        n_rot = rot.shape[0]
        j1, j2 = rng.integers(0, n_rot, n_vox), rng.integers(0, n_rot, n_vox)
        ji = rng.integers(0, iso.shape[0], n_vox)
        sig = (1 - f_iso)[:, None] * (w1[:, None] * rot[j1, k, :] + (1 - w1)[:, None] * rot[j2, k, :]) \
            + f_iso[:, None] * iso[ji]
    sigma = 1.0 / snr
    m = sig.shape[1]
    re = sig + sigma * rng.standard_normal((n_vox, m))
    im = sigma * rng.standard_normal((n_vox, m))
    y = np.sqrt(re * re + im * im).astype(np.float32)
    return y, dirs


def lut_index_torch(dirs, htable_dev):
    """``lut_index_numpy`` on a CUDA tensor (float64 (n, 3)); ``htable_dev``: the hash table as an int64 CUDA tensor."""
    import torch
    d = torch.where(dirs[:, 1:2] < 0, -dirs, dirs)
    two_pi = 2.0 * math.pi
    i2 = torch.fmod(torch.atan2(d[:, 1], d[:, 0]), two_pi)
    i2 = torch.where(i2 < 0, torch.fmod(i2 + two_pi, two_pi), i2)
    big = i2 > math.pi
    rxy = torch.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)
    i1 = torch.where(big, torch.atan2(rxy, -d[:, 2]), torch.atan2(rxy, d[:, 2]))
    i2 = torch.where(big, torch.fmod(torch.atan2(-d[:, 1], -d[:, 0]), two_pi), i2)
    c_round = lambda x: torch.sign(x) * torch.floor(torch.abs(x) + 0.5)
    ii1 = c_round(i1 / math.pi * 180.0).long()
    ii2 = c_round(i2 / math.pi * 180.0).long()
    return htable_dev[ii1 * 181 + ii2]


def make_voxels_torch(model, K, htable, n_vox, seed, device, snr=30.0, iso_max=0.5, chunk=1 << 20):
    """``make_voxels`` generated ON the GPU in chunks (same recipe: 1-2 atoms + isotropic fraction, Rician noise; torch's
    generator, so not the numpy stream) -- for the volumes of BASELINE cfg3 / cfg5 whose float64 intermediates (10.5 M x 288) do
    not belong on the host.  Returns CUDA tensors ``y`` (n_vox, m) float32 and ``dirs`` (n_vox, 3) float64."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if model == "NODDI":
        rot, iso = K["wm"], K["iso"][None, :]
    elif model in ("FreeWater", "FreeWaterMouse"):
        rot, iso = K["D"], K["CSF"]
    elif model == "CylinderZeppelinBall":
        rot, iso = np.concatenate([K["wmr"], K["wmh"]]), K["iso"]
    else:
        raise ValueError(model)
    rot_d = torch.from_numpy(np.ascontiguousarray(rot, dtype=np.float32)).to(device)
    iso_d = torch.from_numpy(np.ascontiguousarray(iso, dtype=np.float32)).to(device)
    ht = torch.from_numpy(np.asarray(htable).astype(np.int64)).to(device)
    n_rot, _, m = rot_d.shape
    y = torch.empty((n_vox, m), dtype=torch.float32, device=device)
    dirs = torch.empty((n_vox, 3), dtype=torch.float64, device=device)
    sigma = 1.0 / snr
    for o in range(0, n_vox, chunk):
        c = min(chunk, n_vox - o)
        v = torch.randn((c, 3), generator=g, device=device, dtype=torch.float64)
        d = v / v.norm(dim=1, keepdim=True)
        k = lut_index_torch(d, ht)
        u = torch.rand((c, 3), generator=g, device=device, dtype=torch.float32)
        f_iso = (u[:, 0] * iso_max)[:, None]
        w1 = torch.where(u[:, 1] < 0.5, 0.3 + 0.4 * u[:, 2], torch.ones_like(u[:, 2]))[:, None]
        j1 = torch.randint(0, n_rot, (c,), generator=g, device=device)
        j2 = torch.randint(0, n_rot, (c,), generator=g, device=device)
        ji = torch.randint(0, iso_d.shape[0], (c,), generator=g, device=device)
        sig = (1 - f_iso) * (w1 * rot_d[j1, k] + (1 - w1) * rot_d[j2, k]) + f_iso * iso_d[ji]
        re = sig + sigma * torch.randn((c, m), generator=g, device=device, dtype=torch.float32)
        im = sigma * torch.randn((c, m), generator=g, device=device, dtype=torch.float32)
        y[o:o + c] = torch.sqrt(re * re + im * im)
        dirs[o:o + c] = d
    return y, dirs


def make_sandi_raw_torch(scheme_full, n_vox, seed, device, snr=30.0, chunk=1 << 19):
    """Raw (un-normalised, un-averaged) SANDI volume (n_vox, nS) float32 on the GPU: the isotropic bi-exponential signal of
    ``make_raw_volume`` times S0 ~ U(400, 1600), Rician noise."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    b = torch.from_numpy(np.asarray(scheme_full.b, dtype=np.float32) * 1e-3).to(device)[None, :]
    nS = b.shape[1]
    out = torch.empty((n_vox, nS), dtype=torch.float32, device=device)
    sigma = 1.0 / snr
    for o in range(0, n_vox, chunk):
        c = min(chunk, n_vox - o)
        u = torch.rand((c, 3), generator=g, device=device, dtype=torch.float32)
        d, f, s0 = 0.3 + 2.2 * u[:, 0:1], 0.2 + 0.6 * u[:, 1:2], 400.0 + 1200.0 * u[:, 2:3]
        sig = f * torch.exp(-b * d) + (1 - f) * torch.exp(-b * 0.2 * d)
        re = sig + sigma * torch.randn((c, nS), generator=g, device=device, dtype=torch.float32)
        im = sigma * torch.randn((c, nS), generator=g, device=device, dtype=torch.float32)
        out[o:o + c] = torch.sqrt(re * re + im * im) * s0
    return out


@dataclass
class Problem:
    """Everything ``model.fit(evaluation)`` reads, for one synthetic config."""
    cfg: int
    model: str
    scheme: Scheme
    lut_dirs: np.ndarray
    htable: np.ndarray
    KERNELS: dict
    params: dict
    y: np.ndarray = field(default=None)
    DIRs: np.ndarray = field(default=None)


def make_problem(cfg, n_vox=None, ndirs=500, model=None, seed=None, snr=30.0, lut_dirs=None, htable=None):
    """``lut_dirs`` / ``htable``: use this direction set and hash table (e.g. the reference's own ``directions/*.bin``)
    instead of the synthetic half-sphere set."""
    model = model or CONFIGS[cfg][0]
    scheme = make_scheme(cfg)
    if model == "SANDI":
        scheme = directional_average_scheme(scheme)
    lut = lut_directions(ndirs) if lut_dirs is None else np.ascontiguousarray(lut_dirs, dtype=np.float64)
    ht = build_htable(lut) if htable is None else np.ascontiguousarray(htable, dtype=np.int16)
    K, p = make_kernels(model, scheme, lut)
    if n_vox is None:
        n_vox = int(np.prod(CONFIGS[cfg][1]))
    y, dirs = make_voxels(model, K, ht, n_vox, (20251017 + cfg) if seed is None else seed, snr=snr)
    return Problem(cfg, model, scheme, lut, ht, K, p, y, dirs)


# --------------------------------------------------------------------------- raw volumes (pre-processing / DTI tests)
def make_raw_volume(cfg, dims, seed=0, snr=30.0, mask_kind="ellipsoid", model=None, ndirs=500):
    """A raw (un-normalised) float32 4-D volume of protocol ``cfg`` as ``Evaluation.load_data`` would read it.

    Every voxel: S0 ~ U(400, 1600) times a single-fibre signal of the model's dictionary (direction uniform on the
    sphere) with an isotropic fraction, Rician noise.  Returns (P, dwi (X, Y, Z, nS) float32, mask (X, Y, Z) uint8)
    where ``P`` is the ``Problem`` carrying scheme / KERNELS / htable.  SANDI (cfg 4): the un-averaged 192-volume
    protocol with an isotropic multi-exponential signal (the directional average is what is under test).
    """
    rng = np.random.default_rng(seed + 1000 * cfg)
    n_tot = int(np.prod(dims))
    mdl = model or CONFIGS[cfg][0]
    P = make_problem(cfg, n_vox=n_tot, ndirs=ndirs, model=mdl, seed=seed + 77, snr=snr)
    if mdl == "SANDI":
        full = make_scheme(cfg)
        b = full.b * 1e-3
        d = rng.uniform(0.3, 2.5, (n_tot, 1))
        f = rng.uniform(0.2, 0.8, (n_tot, 1))
        sig = f * np.exp(-b[None, :] * d) + (1 - f) * np.exp(-b[None, :] * 0.2 * d)
        sigma = 1.0 / snr
        sig = np.sqrt((sig + sigma * rng.standard_normal(sig.shape)) ** 2 + (sigma * rng.standard_normal(sig.shape)) ** 2)
        P.full_scheme = full
    else:
        sig = P.y.astype(np.float64)
        P.full_scheme = P.scheme
    s0 = rng.uniform(400.0, 1600.0, (n_tot, 1))
    dwi = (sig * s0).astype(np.float32).reshape(tuple(dims) + (sig.shape[1],))
    if mask_kind == "ones":
        mask = np.ones(dims, dtype=np.uint8)
    else:
        g = np.meshgrid(*[np.linspace(-1, 1, d) for d in dims], indexing="ij")
        mask = ((g[0] ** 2 + g[1] ** 2 + g[2] ** 2) <= 0.9).astype(np.uint8)
    return P, dwi, mask
