"""The compartment signal models (amico_b200/signals.py) and kernel generation, on the CPU.

* against committed outputs of the REFERENCE's own synthesis (amico/synthesis.py imported from oracle/_ref by
  tests/golden/make_golden.py::synthesis_case) on its 500-direction high-resolution scheme;
* generate -> resample (numpy statement of amico/lut.pyx:274-311) reproduces the directly sampled dictionaries.
"""
import os

import numpy as np
import pytest

from amico_b200 import lut, models, signals, synth
from oracle import pipeline as opl

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _profile_on_grid(atom, g, n_shells):
    """(n_shells * 500,) signal of a fibre along z on the high-resolution scheme: shell s occupies rows 500 s .. 500 s + 499."""
    t = g[:, 2].copy()
    if atom.isotropic:
        return np.concatenate([np.full(500, v) for v in atom.per_shell])
    return np.concatenate([f(t) for f in atom.per_shell])


def test_signal_models_match_reference_synthesis():
    G = np.load(os.path.join(GOLDEN, "synthesis_ref.npz"))
    g = G["grad"]
    s2, s5 = synth.make_scheme(2), synth.make_scheme(5)
    for (od, vf), ref in zip(G["noddi_od_vf"], G["noddi"]):
        a = signals.atom_noddi(s2, 1.7e-3, 1.0 / np.tan(od * np.pi / 2.0), vf)
        assert np.abs(_profile_on_grid(a, g, 2) - ref).max() < 5e-7, (od, vf)   # lmax = 12 truncation at kappa = 21
    np.testing.assert_allclose(_profile_on_grid(signals.atom_ball(s2, 3.0e-3), g, 2), G["noddi_iso"], rtol=1e-10)
    for R, ref in zip(G["cylinder_R"], G["cylinder"]):
        assert np.abs(_profile_on_grid(signals.atom_cylinder(s5, 0.6e-3, R), g, 4) - ref).max() < 2e-5, R
    assert np.abs(_profile_on_grid(signals.atom_zeppelin(s5, 0.6e-3, 0.51e-3), g, 4) - G["zeppelin"]).max() < 1e-6
    # a bare stick is the sharpest profile: exact at b = 1000 (1e-6), visibly truncated by lmax = 12 at b = 4000 (2e-3) -- the
    # reference truncates the same way when it rotates through its lmax = 12 SH fit (amico/lut.pyx:252-264)
    d = np.abs(_profile_on_grid(signals.atom_stick(s5, 1.7e-3), g, 4) - G["stick"])
    assert d[:500].max() < 5e-6 and d.max() < 3e-3
    np.testing.assert_allclose(_profile_on_grid(signals.atom_ball(s5, 2.0e-3), g, 4), G["ball"], rtol=5e-6)  # b from (G, Delta, delta): rounding of the protocol table
    for R, ref in zip(G["sphere_R"], G["sphere"]):
        np.testing.assert_allclose(_profile_on_grid(signals.atom_sphere(s5, 3.0e-3, R), g, 4), ref, rtol=1e-6, err_msg=str(R))
    np.testing.assert_allclose(_profile_on_grid(signals.atom_astrosticks(s5, 1.5e-3), g, 4), G["astrosticks"], rtol=1e-8)


@pytest.mark.parametrize("cfg,name", [(1, "FreeWater"), (2, "NODDI"), (5, "CylinderZeppelinBall"), (4, "SANDI")])
def test_generate_then_resample_reproduces_the_dictionary(tmp_path, cfg, name):
    sch = synth.make_scheme(cfg)
    if name == "SANDI":
        sch = synth.directional_average_scheme(sch)
    dirs = synth.lut_directions(24)
    mdl = getattr(models, name)()
    if name == "NODDI":
        mdl.set(IC_VFs=np.linspace(0.1, 0.99, 3), IC_ODs=np.array([0.03, 0.5]))
    mdl.scheme = sch
    aux = lut.precompute_rotation_matrices(12, dirs)
    idx_in, idx_out_gen = lut.aux_structures_generate(sch, 12)
    mdl.generate(str(tmp_path), aux, idx_in, idx_out_gen, len(dirs))
    first = np.load(tmp_path / "A_001.npy")
    n_coef = 91 * len(sch.shells)
    assert first.dtype == np.float32 and first.shape == ((n_coef,) if name == "SANDI" else (len(dirs), n_coef))
    idx_out, Ylm = lut.aux_structures_resample(sch, 12)
    K = opl.model_resample(name, mdl.get_params(), sch, str(tmp_path), idx_out, Ylm, False, len(dirs), kernel=opl.resample_kernel_exact)
    K0, _ = synth.make_kernels(name, sch, dirs, mdl.get_params() if name == "NODDI" else None)
    for k in K0:
        if k != "model":
            np.testing.assert_allclose(np.asarray(K[k], dtype=np.float64), np.asarray(K0[k], dtype=np.float64), rtol=0, atol=3e-7, err_msg=k)
