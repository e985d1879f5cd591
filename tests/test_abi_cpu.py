"""The C-ABI library loads and exports every symbol include/amico_b200.h declares; without a GPU every
entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from amico_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "amico_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms
    assert lib.amx_version() == 100


def test_fit_args_struct_layout_matches_header():
    src = open(os.path.join(ROOT, "include", "amico_b200.h")).read()
    body = src[src.index("typedef struct amx_fit_args {"):src.index("} amx_fit_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*?\s*\*?([a-z_0-9]+)\s*;", body)
    assert names == [f[0] for f in _lib.FitArgs._fields_]


def test_no_cpu_fallback(gpu_available):
    if gpu_available:
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert lib.amx_device_count() <= 0
    from amico_b200.plan import Plan
    P = synth.make_problem(1, n_vox=8)
    with pytest.raises(_lib.AmxError, match="no CPU fallback"):
        Plan("FreeWater", P.KERNELS, P.htable, P.params)
    from amico_b200 import models

    class Ev:
        y, DIRs, htable, KERNELS, nthreads = P.y.astype(np.float64), np.array(P.DIRs), P.htable, P.KERNELS, 1

        def get_config(self, k):
            return False

    m = models.FreeWater()
    m.set_solver()
    with pytest.raises(RuntimeError):
        m.fit(Ev())


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.amx_plan_create_freewater(0, 0, 500, 10, None, 1, None, 0, None, C.byref(h)) == _lib.AMX_E_INVALID
    assert b"m and ndirs" in lib.amx_last_error()
    assert lib.amx_fit(None, None, None) == _lib.AMX_E_INVALID
    assert lib.amx_plan_destroy(None) == 0
