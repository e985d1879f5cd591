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
    return os.path.isdir(os.path.join(_REF, "amico")) and any(
        f.startswith("models.") and f.endswith(".so") for f in os.listdir(os.path.join(_REF, "amico")))


def _models():
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
