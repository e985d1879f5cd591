"""GPU parity of the callers either side of the fit against oracle/pipeline.py.

Bars: pre-processing (float32 arithmetic, integer compaction) and the result scatter BIT-EXACT; DTI principal directions
(float64, sign-free) within 1e-9 and LUT-index equality >= 99.9 % (a 1-degree LUT cell can flip on a 1e-15 difference);
the whole load -> directions -> fit -> maps flow within the fit's own tolerance.
"""
import ctypes as C

import numpy as np
import pytest

from amico_b200 import _lib as L
from amico_b200 import synth
from amico_b200.evaluation import Evaluation, dti_design_matrix

pytestmark = pytest.mark.gpu


def opl():
    from oracle import pipeline
    return pipeline


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def gpu_preprocess(dwi, sch, mask, *, normalize=True, merge=False, diravg=False, thr=0.0, replace=None, want_b0=True):
    """Host-space amx_preprocess call (the library stages the copies)."""
    lib = L.load()
    dims = dwi.shape[:3]
    n_total = int(np.prod(dims))
    flat = np.ascontiguousarray(dwi.reshape(n_total, -1), dtype=np.float32)
    shells = sorted(sch.shells, key=lambda s: s["b"])
    sh_idx = _i32(np.concatenate([s["idx"] for s in shells]))
    sh_off = _i32(np.concatenate([[0], np.cumsum([len(s["idx"]) for s in shells])]))
    flags = (L.PRE_NORMALIZE if normalize else 0) | (L.PRE_MERGE_B0 if merge else 0) | (L.PRE_DIR_AVG if diravg else 0) \
        | (L.PRE_REPLACE_BAD if replace is not None else 0)
    m_out = 1 + sch.dwi_count if merge else 1 + len(shells) if diravg else sch.nS
    m8 = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(np.uint8).reshape(-1))
    cap = n_total if m8 is None else int((m8 == 1).sum())
    y = np.full((max(cap, 1), m_out), -7.0, dtype=np.float32)
    vidx = np.full(max(cap, 1), -1, dtype=np.int32)
    mb = np.zeros(n_total, dtype=np.float32) if want_b0 else None
    b0, dw = _i32(sch.b0_idx), _i32(sch.dwi_idx)
    a = L.PreArgs()
    a.space, a.device, a.dwi, a.n_total, a.nS = L.SPACE_HOST, 0, flat.ctypes.data, n_total, sch.nS
    a.mask = None if m8 is None else m8.ctypes.data
    a.b0_idx, a.b0_count, a.dwi_idx, a.dwi_count = b0.ctypes.data, len(b0), dw.ctypes.data, len(dw)
    a.shell_idx, a.shell_off, a.n_shells = sh_idx.ctypes.data, sh_off.ctypes.data, len(shells)
    a.flags, a.b0_threshold, a.replace_bad = flags, thr, 0.0 if replace is None else replace
    a.y, a.y_capacity, a.vox_idx = y.ctypes.data, cap, vidx.ctypes.data
    a.mean_b0s = None if mb is None else mb.ctypes.data
    kept, mo = C.c_int64(-1), C.c_int(-1)
    rc = lib.amx_preprocess(C.byref(a), C.byref(kept), C.byref(mo))
    return rc, y[:max(kept.value, 0)], vidx[:max(kept.value, 0)], mb, kept.value, mo.value


@pytest.mark.parametrize("cfg,dims,kw", [
    (1, (8, 8, 8), {}),
    (2, (7, 5, 9), {}),                       # n_total not a multiple of 32
    (2, (7, 5, 9), {"merge": True}),
    (2, (16, 16, 6), {"normalize": False}),
    (5, (6, 6, 5), {"merge": True}),
    (4, (9, 7, 6), {"diravg": True}),
    (4, (9, 7, 6), {"diravg": True, "normalize": False}),
])
@pytest.mark.parametrize("mask_kind", ["ellipsoid", "ones", "none"])
def test_preprocess_bit_exact(cfg, dims, kw, mask_kind):
    P, dwi, mask = synth.make_raw_volume(cfg, dims, seed=cfg, mask_kind="ones" if mask_kind == "none" else mask_kind)
    if mask_kind == "none":
        mask = None
    sch = P.full_scheme
    ref = opl().preprocess(dwi, sch, mask, doNormalizeSignal=kw.get("normalize", True), doMergeB0=kw.get("merge", False),
                           doDirectionalAverage=kw.get("diravg", False))
    rc, y, vidx, mb, kept, m_out = gpu_preprocess(dwi, sch, mask, **kw)
    assert rc == 0, L.load().amx_last_error()
    assert kept == len(ref["vox_idx"]) and m_out == ref["y"].shape[1]
    np.testing.assert_array_equal(vidx, ref["vox_idx"])
    np.testing.assert_array_equal(y.astype(np.float64), ref["y"])
    if ref["mean_b0s"] is not None:
        np.testing.assert_array_equal(mb, ref["mean_b0s"].ravel())


def test_preprocess_mask_values_and_empty():
    P, dwi, mask = synth.make_raw_volume(1, (5, 4, 3), seed=3)
    sch = P.full_scheme
    m = (np.arange(60).reshape(5, 4, 3) % 4).astype(np.uint8)  # values 0..3: only == 1 is kept (core.py:451)
    ref = opl().preprocess(dwi, sch, m)
    rc, y, vidx, _, kept, _ = gpu_preprocess(dwi, sch, m)
    assert rc == 0 and kept == 15
    np.testing.assert_array_equal(vidx, ref["vox_idx"])
    np.testing.assert_array_equal(y.astype(np.float64), ref["y"])
    rc, y, vidx, _, kept, _ = gpu_preprocess(dwi, sch, np.zeros((5, 4, 3), np.uint8))
    assert rc == 0 and kept == 0


def test_preprocess_b0_threshold_and_zero_b0():
    P, dwi, mask = synth.make_raw_volume(2, (6, 6, 6), seed=5)
    sch = P.full_scheme
    dwi = dwi.copy()
    dwi[0, 0, :3] = 0.0                      # dead voxels: mean b0 = 0 -> norm factor 0, not Inf
    dwi[1, 1, 1, sch.b0_idx] *= 1e-3         # weak b0: cropped by b0_min_signal
    ref = opl().preprocess(dwi, sch, mask * 0 + 1, b0_min_signal=0.05)
    rc, y, vidx, mb, kept, _ = gpu_preprocess(dwi, sch, mask * 0 + 1, thr=float(ref["b0_threshold"]))
    assert rc == 0
    np.testing.assert_array_equal(y.astype(np.float64), ref["y"])
    assert (y[:3] == 0).all() and (y[1 * 36 + 1 * 6 + 1] == 0).all()
    # amx_mean_b0 = first half of the normalisation, lets the host form the threshold exactly like core.py:216
    lib = L.load()
    flat = np.ascontiguousarray(dwi.reshape(-1, sch.nS))
    out = np.zeros(len(flat), np.float32)
    b0 = _i32(sch.b0_idx)
    assert lib.amx_mean_b0(L.SPACE_HOST, 0, flat.ctypes.data, len(flat), sch.nS, b0.ctypes.data, len(b0), out.ctypes.data, None) == 0
    np.testing.assert_array_equal(out, ref["mean_b0s"].ravel())
    thr = np.float32(0.05 * out[out > 0].mean())
    assert thr == ref["b0_threshold"]


def test_preprocess_nonfinite_policy():
    P, dwi, mask = synth.make_raw_volume(1, (4, 4, 4), seed=9)
    sch = P.full_scheme
    bad = dwi.copy()
    bad[1, 2, 3, 5] = np.nan
    bad[0, 0, 1, 2] = np.inf
    rc, *_ = gpu_preprocess(bad, sch, mask)
    assert rc == L.AMX_E_NONFINITE and b"Nan or Inf values in the raw signal" in L.load().amx_last_error()
    ref = opl().preprocess(bad, sch, mask * 0 + 1, replace_bad_voxels=0.0)
    rc, y, vidx, _, kept, _ = gpu_preprocess(bad, sch, mask * 0 + 1, replace=0.0)
    assert rc == 0
    np.testing.assert_array_equal(y.astype(np.float64), ref["y"])


def test_preprocess_argument_errors():
    P, dwi, mask = synth.make_raw_volume(1, (4, 4, 4), seed=9)
    rc, *_ = gpu_preprocess(dwi, P.full_scheme, mask, merge=True, diravg=True)
    assert rc == L.AMX_E_INVALID


# ----------------------------------------------------------------------------------------------- DTI directions
def gpu_dti(y, sch, merged=False):
    lib = L.load()
    bvals, bvecs = opl().scheme_gradients(sch, merged)
    W = np.ascontiguousarray(np.linalg.pinv(dti_design_matrix(bvals, bvecs))[:6])
    y = np.ascontiguousarray(y)
    dirs = np.zeros((len(y), 3))
    rc = lib.amx_dti_directions(0, L.SPACE_HOST, y.ctypes.data, L.F64 if y.dtype == np.float64 else L.F32, len(y), y.shape[1],
                                W.ctypes.data, 1e-4, dirs.ctypes.data, None)
    assert rc == 0, lib.amx_last_error()
    return dirs


def gpu_dti_wls(y, sch):
    lib = L.load()
    bvals, bvecs = opl().scheme_gradients(sch, False)
    X = np.ascontiguousarray(dti_design_matrix(bvals, bvecs))
    W = np.ascontiguousarray(np.linalg.pinv(X))
    y = np.ascontiguousarray(y)
    dirs = np.zeros((len(y), 3))
    rc = lib.amx_dti_directions_wls(0, L.SPACE_HOST, y.ctypes.data, L.F64 if y.dtype == np.float64 else L.F32, len(y), y.shape[1],
                                    W.ctypes.data, X.ctypes.data, 1e-4, dirs.ctypes.data, None)
    assert rc == 0, lib.amx_last_error()
    return dirs


@pytest.mark.parametrize("cfg,n,dtype", [(2, 4000, np.float32), (1, 777, np.float64)])
def test_dti_wls_directions_match_oracle(cfg, n, dtype):
    """DTI_fit_method='WLS' (dipy wls_fit_tensor restated in oracle/pipeline.py): the GPU solves the weighted normal equations
    where dipy takes pinv -- same minimiser; directions agree to 1e-8, LUT indices on >= 99.9 % of the voxels."""
    P = synth.make_problem(cfg, n_vox=n)
    y = P.y.astype(dtype)
    ref = opl().dti_directions(y, P.scheme, method="WLS")
    got = gpu_dti_wls(y, P.scheme)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-12)
    dev = 1.0 - np.abs((got * ref).sum(1))
    assert np.percentile(dev, 99) < 1e-10 and dev.max() < 1e-6, (np.percentile(dev, 99), dev.max())
    a = synth.lut_index_numpy(got.copy(), P.htable)
    b = synth.lut_index_numpy(ref.copy(), P.htable)
    assert (a == b).mean() >= 0.999
    ols = opl().dti_directions(y, P.scheme)
    assert (1.0 - np.abs((got * ols).sum(1))).max() > 1e-9  # and it is not the OLS answer


@pytest.mark.parametrize("cfg,n,dtype", [(2, 20000, np.float32), (1, 777, np.float64), (5, 3000, np.float32)])
def test_dti_directions_match_oracle(cfg, n, dtype):
    P = synth.make_problem(cfg, n_vox=n)
    y = P.y.astype(dtype)
    ref = opl().dti_directions(y, P.scheme)
    got = gpu_dti(y, P.scheme)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-12)
    dev = 1.0 - np.abs((got * ref).sum(1))
    assert np.percentile(dev, 99) < 1e-12 and dev.max() < 1e-9, (np.percentile(dev, 99), dev.max())
    ht = P.htable
    a = synth.lut_index_numpy(got.copy(), ht)
    b = synth.lut_index_numpy(ref.copy(), ht)
    assert (a == b).mean() >= 0.999


def test_dti_degenerate_inputs_are_finite():
    sch = synth.make_scheme(2)
    y = np.zeros((40, sch.nS), np.float32)
    y[20:] = 1.0
    d = gpu_dti(y, sch)
    assert np.isfinite(d).all()
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-12)


# ----------------------------------------------------------------------------------------------- result scatter
def test_scatter_maps_bit_exact():
    lib = L.load()
    rng = np.random.default_rng(1)
    n_total, n_vox, k = 5000, 1234, 3
    idx = _i32(np.sort(rng.choice(n_total, n_vox, replace=False)))
    vals = rng.standard_normal((n_vox, k))
    vol = np.full((n_total, k), 9.0, dtype=np.float32)
    assert lib.amx_scatter_maps(0, L.SPACE_HOST, vals.ctypes.data, n_vox, k, idx.ctypes.data, vol.ctypes.data, n_total, None) == 0
    np.testing.assert_array_equal(vol, opl().scatter_maps(vals, idx, n_total))


# ----------------------------------------------------------------------------------------------- whole flow
def _oracle_flow(P, dwi, mask, model, **pre):
    from oracle import oracle as orc
    r = opl().preprocess(dwi, P.full_scheme, mask, **pre)
    dirs = opl().dti_directions(r["y"], P.full_scheme) if model != "SANDI" else None
    l1, l2 = orc.DEFAULT_LAMBDAS[model]
    fit = orc.fit(model, r["y"], None if dirs is None else dirs.copy(), P.htable, P.KERNELS, P.params, l1, l2,
                  dwi_idx=P.scheme.dwi_idx)
    n_total = int(np.prod(dwi.shape[:3]))
    return r, dirs, opl().scatter_maps(fit["estimates"], r["vox_idx"], n_total).reshape(dwi.shape[:3] + (-1,))


@pytest.mark.parametrize("cfg,model,dims", [(1, "FreeWater", (8, 8, 8)), (2, "NODDI", (12, 12, 10)), (5, "CylinderZeppelinBall", (8, 8, 6))])
def test_evaluation_flow_matches_oracle(cfg, model, dims):
    P, dwi, mask = synth.make_raw_volume(cfg, dims, seed=11)
    r, dirs, maps_ref = _oracle_flow(P, dwi, mask, model)
    ae = Evaluation()
    ae.load_data(dwi, P.full_scheme, mask)
    ae.set_model(model)
    ae.load_kernels(P.KERNELS, P.htable)
    res = ae.fit()
    np.testing.assert_array_equal(ae.y, r["y"])
    d = ae.DIRs
    assert (1.0 - np.abs((d * dirs).sum(1))).max() < 1e-9
    maps = res["MAPs"]
    assert maps.dtype == np.float32 and maps.shape == maps_ref.shape
    assert (maps[mask != 1] == 0).all()
    rel = np.abs(maps - maps_ref) / np.maximum(np.abs(maps_ref), 1e-3)
    frac = float((rel[mask == 1] <= 1e-4).all(axis=1).mean())
    # the direction's LUT cell can differ for a voxel whose angle sits on a 1-degree boundary
    assert frac >= 0.995, frac
    assert res["DIRs"].shape == dims + (3,)
    assert ae.get_config("fit_time") > 0


def test_evaluation_flow_sandi_directional_average():
    P, dwi, mask = synth.make_raw_volume(4, (8, 7, 6), seed=2)
    r, _, maps_ref = _oracle_flow(P, dwi, mask, "SANDI", doDirectionalAverage=True)
    ae = Evaluation()
    ae.set_config("doDirectionalAverage", True)
    ae.load_data(dwi, P.full_scheme, mask)
    ae.set_model("SANDI")
    ae.load_kernels(P.KERNELS)
    res = ae.fit()
    np.testing.assert_array_equal(ae.y, r["y"])
    assert ae.scheme.nS == 4
    np.testing.assert_allclose(res["MAPs"], maps_ref, rtol=1e-5, atol=1e-7)
    ae.set_config("amx_exact", True)  # bit-reproducible kernels: float32 cast of equal float64 maps
    np.testing.assert_array_equal(ae.fit()["MAPs"], maps_ref)


def test_file_based_flow_and_save_results(tmp_path):
    """DWI.nii.gz + DWI.scheme + mask.nii.gz in a subject folder -> load_data(file names) -> fit -> save_results: the same
    maps as the array-based flow, written under AMICO/<model>/ with the reference's file names (core.py:501-648)."""
    import os
    from amico_b200 import nifti
    P, dwi, mask = synth.make_raw_volume(1, (6, 7, 5), seed=21)
    subj = tmp_path / "study" / "subj"
    os.makedirs(subj)
    A = np.diag([2.0, 2.0, 2.5, 1.0])
    nifti.save(subj / "DWI.nii.gz", dwi, affine=A)
    nifti.save(subj / "mask.nii", mask.astype(np.float32), affine=A)
    np.savetxt(subj / "DWI.scheme", P.full_scheme.raw, fmt="%.8f", header="VERSION: BVECTOR", comments="")
    ae = Evaluation(str(tmp_path / "study"), "subj")
    ae.set_config("doComputeRMSE", True)
    ae.load_data("DWI.nii.gz", "DWI.scheme", "mask.nii")
    ae.set_model("FreeWater")
    ae.load_kernels(P.KERNELS, P.htable)
    res = ae.fit()
    ref = Evaluation()
    ref.set_config("doComputeRMSE", True)
    ref.load_data(dwi, P.full_scheme.raw, mask)
    ref.set_model("FreeWater")
    ref.load_kernels(P.KERNELS, P.htable)
    want = ref.fit()
    np.testing.assert_array_equal(res["MAPs"], want["MAPs"])
    np.testing.assert_array_equal(res["RMSE"], want["RMSE"])
    np.testing.assert_array_equal(ae.y, ref.y)          # rows come back in the reference's C-scan order
    np.testing.assert_array_equal(ae.mean_b0s, ref.mean_b0s)
    out = ae.save_results()
    assert out.endswith(os.path.join("AMICO", "FreeWater"))
    names = sorted(os.listdir(out))
    assert names == ["config.pickle", "fit_FW.nii.gz", "fit_FiberVolume.nii.gz", "fit_RMSE.nii.gz", "fit_dir.nii.gz"]
    fw = nifti.load(os.path.join(out, "fit_FW.nii.gz"))
    np.testing.assert_array_equal(fw.data, res["MAPs"][..., 1])
    np.testing.assert_allclose(fw.affine, A)
    assert fw.header[148:228].startswith(b"Isotropic free-water volume fraction (AMICO v")
    d = nifti.load(os.path.join(out, "fit_dir.nii.gz"))
    assert d.shape == (6, 7, 5, 3)


@pytest.mark.parametrize("dtype,code,slope,inter", [(np.float32, 16, 1.0, 0.0), (np.int16, 4, 0.37, 11.5), (np.uint8, 2, float("nan"), 0.0),
                                                    (np.uint16, 512, 2.0, 0.0), (np.float64, 64, 1.0, 0.0), (np.int32, 8, 0.0, 5.0)])
def test_volume_to_voxel_major_matches_numpy(dtype, code, slope, inter):
    """On-disk block [nS][n_total] of any NIfTI dtype -> float32 [n_total][nS], scaling like nibabel's get_fdata + astype."""
    lib = L.load()
    rng = np.random.default_rng(3)
    n_total, nS = 1000, 37      # neither a multiple of the 32 x 32 tile
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    src = (rng.integers(max(info.min, -3000), min(info.max, 3000), (nS, n_total)).astype(dtype) if info
           else (rng.standard_normal((nS, n_total)) * 100).astype(dtype))
    dst = np.zeros((n_total, nS), np.float32)
    rc = lib.amx_volume_to_voxel_major(0, L.SPACE_HOST, src.ctypes.data, code, n_total, nS, slope, inter, dst.ctypes.data, None)
    assert rc == 0, lib.amx_last_error()
    f = src.astype(np.float64)
    if np.isfinite(slope) and np.isfinite(inter) and slope != 0 and (slope != 1 or inter != 0):
        f = f * slope + inter
    np.testing.assert_array_equal(dst, f.astype(np.float32).T)


def test_file_based_flow_int16_scaled(tmp_path):
    """An int16 image with scl_slope / scl_inter goes through the same flow as its float32 array."""
    import os
    import struct
    from amico_b200 import nifti
    P, dwi, mask = synth.make_raw_volume(1, (5, 6, 4), seed=8)
    q = np.round(dwi / 0.25).astype(np.int16)            # quantised: raw * 0.25 reproduces a float32-exact volume
    vol = (q.astype(np.float64) * 0.25).astype(np.float32)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 4, 5, 6, 4, q.shape[3], 1, 1, 1)
    struct.pack_into("<h", hdr, 70, 4)
    struct.pack_into("<h", hdr, 72, 16)
    struct.pack_into("<8f", hdr, 76, 1, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<f", hdr, 112, 0.25)
    hdr[344:348] = b"n+1\0"
    os.makedirs(tmp_path / "s")
    with open(tmp_path / "s" / "DWI.nii", "wb") as f:
        f.write(bytes(hdr) + bytes(4) + q.tobytes(order="F"))
    ae = Evaluation(str(tmp_path), "s")
    ae.load_data("DWI.nii", P.full_scheme.raw, mask)
    ref = Evaluation()
    ref.load_data(vol, P.full_scheme.raw, mask)
    np.testing.assert_array_equal(ae.y, ref.y)


SHARD_WORKER = """
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from amico_b200 import synth
from amico_b200.evaluation import Evaluation
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P, dwi, mask = synth.make_raw_volume(1, (9, 7, 5), seed=13)
ae = Evaluation(device=0)                       # both ranks share GPU 0 in this test; one GPU per rank in production
ae.set_config("doComputeRMSE", True)
ae.load_data(dwi, P.full_scheme, mask, b0_min_signal=0.01, shard=(rank, world))
ae.set_model("FreeWater")
ae.load_kernels(P.KERNELS, P.htable)
res = ae.fit()
if rank == 0:
    one = Evaluation(device=0)
    one.set_config("doComputeRMSE", True)
    one.load_data(dwi, P.full_scheme, mask, b0_min_signal=0.01)
    one.set_model("FreeWater")
    one.load_kernels(P.KERNELS, P.htable)
    want = one.fit()
    for k in ("MAPs", "DIRs", "RMSE"):
        assert res[k].shape == want[k].shape and np.array_equal(res[k], want[k]), k
else:
    assert res is None
dist.barrier()
dist.destroy_process_group()
print("worker", rank, "ok")
"""


def test_sharded_evaluation_world2(tmp_path):
    """Two ranks (gloo, sharing the one test GPU) each run the flow on their voxel slab; rank 0 gathers the full volumes,
    identical to the single-process result."""
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER % root)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


def test_evaluation_flow_noddi_merged_b0():
    """doMergeB0: [mean b0 | dwi] volume (core.py:224-227), DTI on the merged gradient table (core.py:432-433), NODDI on the
    merged KERNELS (m == 1 + dwi_count -> rows 1..m-1 are the DWI rows, amico/models.pyx:916-918)."""
    from oracle import oracle as orc
    P, dwi, mask = synth.make_raw_volume(2, (10, 9, 8), seed=17)
    sch = P.full_scheme
    merge_idx = np.hstack((sch.b0_idx[0], sch.dwi_idx))
    K = dict(P.KERNELS)
    K["wm"] = np.ascontiguousarray(P.KERNELS["wm"][:, :, merge_idx])
    K["iso"] = np.ascontiguousarray(P.KERNELS["iso"][merge_idx])
    r = opl().preprocess(dwi, sch, mask, doMergeB0=True)
    dirs = opl().dti_directions(r["y"], sch, doMergeB0=True)
    l1, l2 = orc.DEFAULT_LAMBDAS["NODDI"]
    ref = orc.fit("NODDI", r["y"], dirs.copy(), P.htable, K, P.params, l1, l2, dwi_idx=sch.dwi_idx)
    maps_ref = opl().scatter_maps(ref["estimates"], r["vox_idx"], mask.size).reshape(mask.shape + (-1,))
    ae = Evaluation()
    ae.set_config("doMergeB0", True)
    ae.load_data(dwi, sch, mask)
    ae.set_model("NODDI")
    ae.load_kernels(K, P.htable)
    res = ae.fit()
    np.testing.assert_array_equal(ae.y, r["y"])
    assert ae.y.shape[1] == 1 + sch.dwi_count
    rel = np.abs(res["MAPs"] - maps_ref) / np.maximum(np.abs(maps_ref), 1e-3)
    assert float((rel[mask == 1] <= 1e-4).all(axis=1).mean()) >= 0.995


def test_reference_style_script_end_to_end(tmp_path):
    """The reference's usage pattern, start to finish, without the reference: files in a study folder -> load_data ->
    set_model -> generate_kernels -> load_kernels -> fit -> save_results; maps equal to the flow fed with the directly
    sampled dictionary."""
    import os
    from amico_b200 import nifti
    P, dwi, mask = synth.make_raw_volume(1, (6, 6, 6), seed=31)
    subj = tmp_path / "study" / "s01"
    os.makedirs(subj)
    nifti.save(subj / "DWI.nii.gz", dwi)
    nifti.save(subj / "mask.nii.gz", mask.astype(np.float32))
    np.savetxt(subj / "DWI.scheme", P.full_scheme.raw, fmt="%.8f", header="VERSION: BVECTOR", comments="")
    ae = Evaluation(str(tmp_path / "study"), "s01")
    ae.load_data("DWI.nii.gz", "DWI.scheme", "mask.nii.gz", b0_thr=0)
    ae.set_model("FreeWater")
    ae.set_lut(P.lut_dirs, P.htable)
    path = ae.generate_kernels(regenerate=True)
    assert len(os.listdir(path)) == 11 and ae.generate_kernels() == path          # second call: kernels are reused
    ae.load_kernels()
    res = ae.fit()
    ae.save_results()
    ref = Evaluation()
    ref.load_data(dwi, P.full_scheme.raw, mask)
    ref.set_model("FreeWater")
    ref.load_kernels(P.KERNELS, P.htable)
    want = ref.fit()
    assert np.abs(res["MAPs"] - want["MAPs"]).max() < 1e-4
    assert os.path.isfile(subj / "AMICO" / "FreeWater" / "fit_FW.nii.gz")
