"""Run the reference's own ``<Model>.fit(evaluation)`` from ``oracle/_ref`` (see ``build_ref.py``).

TEST INFRASTRUCTURE ONLY.  ``amico.models`` here is daducci/AMICO's Cython compiled unmodified, with the
absent spams-cython solvers bound to ``amico_oracle.c`` -- "reference glue + restated solvers".
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    d = os.path.join(_REF, "amico")
    need = ("models.", "lut.", "util.", "scheme.", "synthesis.")
    return os.path.isdir(d) and all(any(f.startswith(n) and f.endswith(".so") for f in os.listdir(d)) for n in need)


class _ProgressBar:
    """No-op stand-in for dicelib.ui.ProgressBar (call shapes: amico/models.pyx:304-310, 802; core.py:457)."""

    def __init__(self, total=None, multithread_progress=None, disable=False, **kw):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a, **kw):
        pass


def _stub(*a, **k):
    raise NotImplementedError("dipy stub (oracle/_ref): only kernel generation calls this, which is outside the hot path")


def _install_stubs():
    """The compiled reference modules import dicelib.ui.ProgressBar and three dipy symbols at import time; neither package is
    installed here.  The stand-ins are created in sys.modules (not as files), so that oracle/_ref needs nothing but its
    extension modules to be usable on a box that received only the built binaries."""
    import types

    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        if "." in name:
            setattr(sys.modules[name.rsplit(".", 1)[0]], name.rsplit(".", 1)[1], m)
        sys.modules[name] = m
        return m

    try:
        import dicelib.ui  # noqa: F401
    except Exception:
        mod("dicelib", __path__=[])
        mod("dicelib.ui", ProgressBar=_ProgressBar)
    try:
        import dipy.reconst.shm  # noqa: F401
        import dipy.core.geometry  # noqa: F401
        import dipy.data.fetcher  # noqa: F401
    except Exception:
        for k in [k for k in sys.modules if k == "dipy" or k.startswith("dipy.")]:
            del sys.modules[k]
        mod("dipy", __path__=[])
        mod("dipy.data", __path__=[])
        mod("dipy.data.fetcher", dipy_home=os.path.join(os.path.expanduser("~"), ".dipy"))
        mod("dipy.core", __path__=[])
        mod("dipy.core.geometry", cart2sphere=_stub)
        mod("dipy.reconst", __path__=[])
        mod("dipy.reconst.shm", real_sh_descoteaux=_stub)
    if "amico" not in sys.modules and not os.path.exists(os.path.join(_REF, "amico", "__init__.py")):
        mod("amico", __path__=[os.path.join(_REF, "amico")])


def _models():
    _install_stubs()
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    util = importlib.import_module("amico.util")
    util.set_verbose(1)
    return importlib.import_module("amico.models")


class _Evaluation:
    """The attributes ``fit`` reads from ``amico.core.Evaluation``."""

    def __init__(self, y, DIRs, htable, KERNELS, nthreads, config):
        self.y, self.DIRs, self.htable, self.KERNELS, self.nthreads = y, DIRs, htable, KERNELS, nthreads
        self._cfg = config

    def get_config(self, key):
        return self._cfg.get(key)


def make_model(name, scheme, params=None, lambda1=None, lambda2=None):
    M = _models()
    cls = {"FreeWaterMouse": "FreeWater"}.get(name, name)
    model = getattr(M, cls)()
    if params:
        p = {k: v for k, v in params.items()}
        model.set(**p)
    if cls == "CylinderZeppelinBall":
        model.isExvivo = False  # never set by the reference's own set() (SURVEY 8a quirk v)
    model.scheme = scheme
    kw = {}
    if lambda1 is not None:
        kw["lambda1"] = lambda1
    if lambda2 is not None:
        kw["lambda2"] = lambda2
    model.set_solver(**kw)
    return model


def fit_problem(P, nthreads=1, rmse=False, nrmse=False, extra=False, lambda1=None, lambda2=None):
    model = make_model(P.model, P.scheme, P.params, lambda1, lambda2)
    cfg = {"doComputeRMSE": rmse, "doComputeNRMSE": nrmse, "doSaveModulatedMaps": extra, "doSaveCorrectedDWI": extra}
    y = np.ascontiguousarray(P.y, dtype=np.float64)
    dirs = None if P.DIRs is None else np.array(P.DIRs, dtype=np.float64)
    ev = _Evaluation(y, dirs, P.htable, P.KERNELS, nthreads, cfg)
    return model.fit(ev)
