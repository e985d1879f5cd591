"""Host-side mirror of the reference's plugin surface (amico/models.pyx): names, maps, defaults, kwargs."""
import inspect

import numpy as np
import pytest

from amico_b200 import models, parallel


def test_ids_maps_and_solver_defaults():
    exp = {
        "NODDI": (["NDI", "ODI", "FWF"], (5e-1, 1e-3)),                       # models.pyx:668-671, :721
        "FreeWater": (["FiberVolume", "FW"], (0.0, 1e-3)),                    # :1054, :1077
        "CylinderZeppelinBall": (["v", "a", "d"], (0.0, 4.0)),                # :393, :439
        "SANDI": (["fsoma", "fneurite", "fextra", "Rsoma", "Din", "De"], (0.0, 5e-3)),  # :1361, :1405
    }
    for name, (maps, lam) in exp.items():
        m = getattr(models, name)()
        assert m.id == name and m.maps_name == maps and len(m.maps_descr) == len(maps)
        m.set_solver()
        assert (m.solver_params["lambda1"], m.solver_params["lambda2"]) == lam
        assert m.get_params()["id"] == name
        # Evaluation.set_solver filters kwargs with inspect.signature (core.py:316-325)
        assert set(inspect.signature(m.set_solver).parameters) == {"lambda1", "lambda2"}


def test_default_grids():
    n = models.NODDI()
    assert np.allclose(n.IC_VFs, np.linspace(0.1, 0.99, 12)) and len(n.IC_ODs) == 12 and n.dPar == 1.7e-3 and n.dIso == 3e-3
    n.set(isExvivo=True)
    assert n.maps_name[-1] == "dot" and n.get_params()["isExvivo"]
    f = models.FreeWater()
    assert np.allclose(f.d_perps, np.linspace(0.1, 1.0, 10) * 1e-3) and f.d_isos == [2.5e-3]
    f.set(type="Mouse")
    assert f.maps_name == ["FiberVolume", "FW", "FW_blood", "FW_csf"] and f.d_isos == [1.5e-3, 3e-3]
    c = models.CylinderZeppelinBall()
    assert len(c.Rs) == 21 and c.isExvivo is False and len(c.d_perps) == 4
    s = models.SANDI()
    assert len(s.Rs) == 5 and len(s.d_in) == 5 and len(s.d_isos) == 5 and s.d_is == 3e-3


def test_resample_needs_the_gpu(tmp_path):
    # resample is provided (GPU projection): without a device it must fail loudly, never fall back to the CPU
    from amico_b200 import lut, synth
    sch = synth.make_scheme(1)
    fw = models.FreeWater()
    fw.scheme = sch
    idx_out, Ylm = lut.aux_structures_resample(sch, 12)
    for i in range(11):
        np.save(tmp_path / f"A_{i + 1:03d}.npy", np.zeros((3, Ylm.shape[1]) if i < 10 else (Ylm.shape[1],), np.float32))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            fw.resample(str(tmp_path), idx_out, Ylm, False, 3)


def test_kernels_model_id_is_checked():
    m = models.NODDI()

    class Ev:
        KERNELS = {"model": "FreeWater"}
        htable = None

    with pytest.raises(RuntimeError, match="not created with the same model"):
        m._get_plan(Ev())


def test_shard_bounds_cover_and_follow_reference_chunking():
    for n in (0, 1, 5, 100, 1048576, 10485760):
        for w in (1, 2, 3, 4, 8):
            b = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            if n >= w:  # same chunks as BaseModel.fit (models.pyx:205-211)
                c = n // w
                assert all(b[i] == (i * c, (i + 1) * c) for i in range(w - 1))
