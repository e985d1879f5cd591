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


def refdirs_case():
    """NODDI cfg2 on the reference's REAL 500-direction set + hash table: the fixture carries both tables, test directions
    that include every degree-grid edge case, and the reference fit's maps."""
    dirs, ht = reference_tables(500)
    P = synth.make_problem(2, n_vox=384, seed=77, lut_dirs=dirs, htable=ht)
    res = ref_runner.fit_problem(P, nthreads=2, rmse=True)
    out = {k: np.asarray(v) for k, v in res.items()}
    out.update(lut_dirs=dirs, htable=ht, input_sha256=np.array(input_digest(P)))
    np.savez_compressed(os.path.join(HERE, "noddi_refdirs500.npz"), **out)
    print("noddi_refdirs500", {k: v.shape for k, v in out.items()})


def main():
    if not ref_runner.available():
        raise SystemExit("oracle/_ref is not built: run python oracle/build_ref.py first")
    refdirs_case()
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
