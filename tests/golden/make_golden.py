"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own code.

Run in the development container (needs /root/reference and the build of oracle/_ref, see
oracle/build_ref.py):   python tests/golden/make_golden.py

For every model the inputs are the seeded synthetic problem of amico_b200.synth (regenerated
bit-identically by the tests from the seed) and the outputs are what daducci/AMICO's unmodified
Cython `<Model>.fit(evaluation)` returns (amico/models.pyx, compiled from /root/reference) with its
two third-party solver entry points (spams-cython, absent) bound to oracle/amico_oracle.c.  A
fixture stores the outputs plus a checksum of the inputs so that a drift in the generator is caught.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from amico_b200 import synth  # noqa: E402
from oracle import ref_runner  # noqa: E402

CASES = [
    # name, cfg, model, n_vox, seed, flags
    ("freewater_cfg1", 1, "FreeWater", 512, None, dict(rmse=True, nrmse=True, extra=True)),
    ("freewater_mouse", 1, "FreeWaterMouse", 256, 11, dict(rmse=True, nrmse=True, extra=True)),
    ("noddi_cfg2", 2, "NODDI", 384, None, dict(rmse=True, nrmse=True, extra=True)),
    ("sandi_cfg4", 4, "SANDI", 512, None, dict(rmse=True, nrmse=True)),
    ("czb_cfg5", 5, "CylinderZeppelinBall", 320, None, dict(rmse=True, nrmse=True)),
]


def input_digest(P):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(P.y).tobytes())
    if P.DIRs is not None:
        h.update(np.ascontiguousarray(P.DIRs).tobytes())
    for k in sorted(P.KERNELS):
        if k != "model":
            h.update(np.ascontiguousarray(P.KERNELS[k]).tobytes())
    return h.hexdigest()


def reference_tables(ndirs=500):
    """The reference's own LUT direction set and hash table (amico/directions/*.bin), read the way amico/lut.pyx:50-91 does."""
    d = os.path.join("/root/reference", "amico", "directions")
    dirs = np.fromfile(os.path.join(d, "ndirs=%d.bin" % ndirs), dtype=np.float64).reshape(ndirs, 3)
    ht = np.fromfile(os.path.join(d, "htable_ndirs=%d.bin" % ndirs), dtype=np.int16)
    return dirs, ht


def refdirs_case(ndirs=500, cfg=2, n_vox=384, seed=77):
    """NODDI on the reference's REAL direction sets + hash tables (amico/directions/ndirs=*.bin: 1, 500 ... 10000, 32761 --
    amico/lut.pyx:18-25): the fixture carries both tables, test directions that include every degree-grid edge case, and the
    reference fit's maps.  ndirs = 1 (one atom orientation for every voxel) and a large set exercise the table sizes either
    side of the default 500."""
    dirs, ht = reference_tables(ndirs)
    P = synth.make_problem(cfg, n_vox=n_vox, seed=seed, lut_dirs=dirs, htable=ht)
    res = ref_runner.fit_problem(P, nthreads=2, rmse=True)
    out = {k: np.asarray(v) for k, v in res.items()}
    out.update(lut_dirs=dirs, htable=ht, input_sha256=np.array(input_digest(P)))
    np.savez_compressed(os.path.join(HERE, "noddi_refdirs%d.npz" % ndirs), **out)
    print("noddi_refdirs%d" % ndirs, {k: v.shape for k, v in out.items()})


def synthesis_case():
    """Signals of the reference's own compartment models (amico/synthesis.py, imported from oracle/_ref) for a fibre along z on
    its 500-direction high-resolution scheme (amico/lut.pyx:359-384, gradient table lut.pyx:390-891), one entry per atom type."""
    import importlib
    ref_runner._models()  # puts oracle/_ref on sys.path
    syn = importlib.import_module("amico.synthesis")
    lut = importlib.import_module("amico.lut")
    rsch = importlib.import_module("amico.scheme")
    out = {"grad": np.asarray(lut.grad, dtype=np.float64)}
    sch2 = synth.make_scheme(2)
    hi2 = lut.create_high_resolution_scheme(rsch.Scheme(sch2.raw.copy(), 0))
    ic, ec, iso = syn.NODDIIntraCellular(hi2), syn.NODDIExtraCellular(hi2), syn.NODDIIsotropic(hi2)
    cases = []
    for od, vf in ((0.03, 0.1), (0.3, 0.5), (0.84, 0.99)):
        kappa = 1.0 / np.tan(od * np.pi / 2.0)
        cases.append(vf * ic.get_signal(1.7e-3, kappa) + (1 - vf) * ec.get_signal(1.7e-3, kappa, vf))
    out["noddi_od_vf"] = np.array([(0.03, 0.1), (0.3, 0.5), (0.84, 0.99)])
    out["noddi"] = np.array(cases)
    out["noddi_iso"] = iso.get_signal(3.0e-3)
    sch5 = synth.make_scheme(5)
    hi5 = lut.create_high_resolution_scheme(rsch.Scheme(sch5.raw.copy(), 0))
    out["cylinder_R"] = np.array([0.01e-6, 2.0e-6, 8.0e-6])
    out["cylinder"] = np.array([syn.CylinderGPD(hi5).get_signal(0.6e-3, R) for R in out["cylinder_R"]])
    out["zeppelin"] = syn.Zeppelin(hi5).get_signal(0.6e-3, 0.51e-3)
    out["stick"] = syn.Stick(hi5).get_signal(1.7e-3)
    out["ball"] = syn.Ball(hi5).get_signal(2.0e-3)
    out["sphere_R"] = np.array([1.0e-6, 6.5e-6, 12.0e-6])
    out["sphere"] = np.array([syn.SphereGPD(hi5).get_signal(3.0e-3, R) for R in out["sphere_R"]])
    out["astrosticks"] = syn.Astrosticks(hi5).get_signal(1.5e-3)
    np.savez_compressed(os.path.join(HERE, "synthesis_ref.npz"), **out)
    print("synthesis_ref", {k: np.asarray(v).shape for k, v in out.items()})


def main():
    if not ref_runner.available():
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    for nd in (500, 1, 5000):
        refdirs_case(nd)
    synthesis_case()
    for name, cfg, model, n_vox, seed, flags in CASES:
        P = synth.make_problem(cfg, n_vox=n_vox, model=model, seed=seed)
        # NB the reference sizes its y_est scratch by the CHUNK's voxel count (models.pyx:588, 875, 1210, 1548:
        # np.zeros(y_view.shape[0])) but writes m entries into it (models.pyx:49-51): chunks shorter than m
        # overflow the heap.  Keep every chunk >= m here.
        nthreads = 2 if n_vox // 2 >= P.y.shape[1] else 1
        assert n_vox // nthreads >= P.y.shape[1]
        res = ref_runner.fit_problem(P, nthreads=nthreads, **flags)
        out = {k: np.asarray(v) for k, v in res.items()}
        out["input_sha256"] = np.array(input_digest(P))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
