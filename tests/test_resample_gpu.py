"""GPU parity of kernel resampling: ``<Model>.resample`` on ``A_###.npy`` files in the layout
the reference's ``generate`` writes, against oracle/pipeline.py.

Bars: bit-exact against the float64-accumulated projection rounded once (``resample_kernel_exact``); within a few float32
ulps of the reference's own float32 ``np.dot`` statement (whose summation order is the BLAS build's, not ours).
"""
import os

import numpy as np
import pytest

from amico_b200 import lut, models, synth

pytestmark = pytest.mark.gpu


def opl():
    from oracle import pipeline
    return pipeline


def _write_atoms(tmp, shapes, seed):
    rng = np.random.default_rng(seed)
    for i, shp in enumerate(shapes):
        np.save(os.path.join(tmp, f"A_{i + 1:03d}.npy"), (rng.standard_normal(shp) * 0.3).astype(np.float32))


CASES = [
    ("FreeWater", 1, False, 37),
    ("NODDI", 2, False, 20),
    ("NODDI", 2, True, 20),
    ("CylinderZeppelinBall", 5, False, 11),
    ("SANDI", 4, False, 5),
]


@pytest.mark.parametrize("name,cfg,merge,ndirs", CASES)
def test_model_resample_matches_oracle(tmp_path, name, cfg, merge, ndirs):
    sch = synth.make_scheme(cfg)
    mdl = getattr(models, name)()
    if name == "NODDI":  # a smaller grid keeps the CPU statement-for-statement oracle quick
        mdl.set(IC_VFs=np.linspace(0.1, 0.99, 4), IC_ODs=np.array([0.03, 0.2, 0.7]))
    mdl.scheme = sch
    idx_out, Ylm_out = lut.aux_structures_resample(sch, 12)
    n_coef = Ylm_out.shape[1]
    p = mdl.get_params()
    if name == "NODDI":
        shapes = [(ndirs, n_coef)] * (len(p["IC_ODs"]) * len(p["IC_VFs"])) + [(n_coef,)]
    elif name == "FreeWater":
        shapes = [(ndirs, n_coef)] * len(p["d_perps"]) + [(n_coef,)] * len(p["d_isos"])
    elif name == "CylinderZeppelinBall":
        shapes = [(ndirs, n_coef)] * (len(p["Rs"]) + len(p["d_perps"])) + [(n_coef,)] * len(p["d_isos"])
    else:
        shapes = [(n_coef,)] * (len(p["Rs"]) + len(p["d_in"]) + len(p["d_isos"]))
    _write_atoms(str(tmp_path), shapes, seed=cfg)
    got = mdl.resample(str(tmp_path), idx_out, Ylm_out, merge, ndirs)
    exact = opl().model_resample(name, p, sch, str(tmp_path), idx_out, Ylm_out, merge, ndirs, kernel=opl().resample_kernel_exact)
    ref = opl().model_resample(name, p, sch, str(tmp_path), idx_out, Ylm_out, merge, ndirs)
    assert set(got) == set(ref)
    for k in ref:
        if k == "model":
            assert got[k] == ref[k]
            continue
        g, e, r = np.asarray(got[k]), np.asarray(exact[k]), np.asarray(ref[k])
        assert g.shape == r.shape and g.dtype == r.dtype, k
        if name == "SANDI" or k == "norms":
            np.testing.assert_allclose(g, e, rtol=1e-6, err_msg=k)  # float32 norm of bit-equal inputs, BLAS nrm2 order
        else:
            np.testing.assert_array_equal(g, e, err_msg=k)
        np.testing.assert_allclose(g, r, rtol=0, atol=2e-5, err_msg=k)  # float32 gemv of ~180 terms of size ~0.1


def test_resample_then_fit_runs(tmp_path):
    """KERNELS built by the GPU resample go straight into a plan (layout check of the whole chain)."""
    from amico_b200.plan import Plan
    sch = synth.make_scheme(1)
    mdl = models.FreeWater()
    mdl.scheme = sch
    idx_out, Ylm_out = lut.aux_structures_resample(sch, 12)
    P = synth.make_problem(1, n_vox=64)
    ndirs = 500
    # SH coefficients of smooth positive functions: project the synthetic signal-space kernels back with a least-squares fit
    Yp = np.linalg.pinv(Ylm_out.astype(np.float64))
    for i in range(10):
        np.save(tmp_path / f"A_{i + 1:03d}.npy", (P.KERNELS["D"][i][:, sch.dwi_idx].astype(np.float64) @ Yp.T).astype(np.float32))
    np.save(tmp_path / "A_011.npy", (Yp @ P.KERNELS["CSF"].reshape(-1)[sch.dwi_idx].astype(np.float64)).astype(np.float32))
    K = mdl.resample(str(tmp_path), idx_out, Ylm_out, False, ndirs)
    assert K["D"].shape == P.KERNELS["D"].shape and (K["D"][:, :, sch.b0_idx] == 1).all()
    with Plan("FreeWater", K, P.htable, P.params, dwi_idx=sch.dwi_idx) as plan:
        res = plan.fit(P.y, np.array(P.DIRs, dtype=np.float64), 0.0, 1e-3)
    assert np.isfinite(res["estimates"]).all()


def test_resample_outdated_lut_error(tmp_path):
    sch = synth.make_scheme(1)
    mdl = models.FreeWater()
    mdl.scheme = sch
    idx_out, Ylm_out = lut.aux_structures_resample(sch, 12)
    _write_atoms(str(tmp_path), [(7, Ylm_out.shape[1])] * 10 + [(Ylm_out.shape[1],)], 0)
    with pytest.raises(RuntimeError, match="Outdated LUT"):
        mdl.resample(str(tmp_path), idx_out, Ylm_out, False, 9)


def test_generate_resample_fit_chain(tmp_path):
    """The whole offline chain without the reference: generate (rotated SH atoms on disk) -> resample on the GPU -> the same
    KERNELS as sampling the models directly (1 float32 ulp) -> identical FreeWater maps."""
    from amico_b200.plan import Plan
    P = synth.make_problem(1, n_vox=256)
    sch = P.scheme
    mdl = models.FreeWater()
    mdl.scheme = sch
    aux = lut.precompute_rotation_matrices(12, P.lut_dirs)
    idx_in, idx_out_gen = lut.aux_structures_generate(sch, 12)
    mdl.generate(str(tmp_path), aux, idx_in, idx_out_gen, len(P.lut_dirs))
    idx_out, Ylm_out = lut.aux_structures_resample(sch, 12)
    K = mdl.resample(str(tmp_path), idx_out, Ylm_out, False, len(P.lut_dirs))
    for k in ("D", "CSF"):
        np.testing.assert_allclose(K[k], P.KERNELS[k], rtol=0, atol=2e-7, err_msg=k)
    with Plan("FreeWater", K, P.htable, P.params, dwi_idx=sch.dwi_idx) as plan:
        a = plan.fit(P.y, np.array(P.DIRs), 0.0, 1e-3)["estimates"]
    with Plan("FreeWater", P.KERNELS, P.htable, P.params, dwi_idx=sch.dwi_idx) as plan:
        b = plan.fit(P.y, np.array(P.DIRs), 0.0, 1e-3)["estimates"]
    assert np.abs(a - b).max() < 1e-4
