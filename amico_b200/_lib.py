"""ctypes binding of ``libamico_b200.so`` (C ABI declared in ``include/amico_b200.h``).

The library is built in-tree by :mod:`amico_b200.build`.  There is no CPU fallback: if the shared
library is missing, or no B200 is visible, every fit raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMICO_B200_LIB") or os.path.join(_HERE, "libamico_b200.so")  # override: A/B builds of the same library

AMX_OK, AMX_E_INVALID, AMX_E_CUDA, AMX_E_LUT_RANGE, AMX_E_CAPACITY, AMX_E_NONFINITE = 0, -1, -2, -3, -4, -5
PRE_NORMALIZE, PRE_MERGE_B0, PRE_DIR_AVG, PRE_REPLACE_BAD = 1, 2, 4, 8
MODEL_NODDI, MODEL_FREEWATER, MODEL_CZB, MODEL_SANDI = 0, 1, 2, 3
FLAG_RMSE, FLAG_NRMSE, FLAG_EXTRA, FLAG_EXACT = 1, 2, 4, 8
F32, F64 = 0, 1
SPACE_HOST, SPACE_DEVICE = 0, 1

# every symbol include/amico_b200.h declares
EXPORTS = [
    "amx_last_error", "amx_version", "amx_device_count",
    "amx_plan_create_noddi", "amx_plan_create_freewater", "amx_plan_create_czb", "amx_plan_create_sandi",
    "amx_plan_destroy", "amx_plan_info", "amx_fit", "amx_lut_indices", "amx_plan_last_timing",
    "amx_plan_last_counters",
    "amx_preprocess", "amx_mean_b0", "amx_dti_directions", "amx_dti_directions_wls", "amx_scatter_maps", "amx_resample_kernels", "amx_volume_to_voxel_major",
]


class FitArgs(C.Structure):
    """``amx_fit_args`` of include/amico_b200.h."""
    _fields_ = [
        ("space", C.c_int), ("y_dtype", C.c_int), ("y", C.c_void_p), ("n_vox", C.c_int64), ("dirs", C.c_void_p),
        ("lambda1", C.c_double), ("lambda2", C.c_double), ("flags", C.c_uint32), ("estimates", C.c_void_p),
        ("rmse", C.c_void_p), ("nrmse", C.c_void_p), ("extra", C.c_void_p), ("lut_out", C.c_void_p),
        ("support_out", C.c_void_p), ("coeff_out", C.c_void_p), ("stream", C.c_void_p),
    ]


class PreArgs(C.Structure):
    """``amx_pre_args`` of include/amico_b200.h."""
    _fields_ = [
        ("space", C.c_int), ("device", C.c_int), ("dwi", C.c_void_p), ("n_total", C.c_int64), ("nS", C.c_int),
        ("mask", C.c_void_p), ("b0_idx", C.c_void_p), ("b0_count", C.c_int), ("dwi_idx", C.c_void_p),
        ("dwi_count", C.c_int), ("shell_idx", C.c_void_p), ("shell_off", C.c_void_p), ("n_shells", C.c_int),
        ("flags", C.c_uint32), ("b0_threshold", C.c_float), ("replace_bad", C.c_float), ("y", C.c_void_p),
        ("y_capacity", C.c_int64), ("vox_idx", C.c_void_p), ("mean_b0s", C.c_void_p), ("stream", C.c_void_p),
    ]


_lib = None


class AmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m amico_b200.build` (needs nvcc). "
            "amico_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.amx_last_error.restype = C.c_char_p
    lib.amx_last_error.argtypes = []
    lib.amx_version.restype = i32
    lib.amx_device_count.restype = i32
    lib.amx_plan_create_noddi.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, C.POINTER(vp)]
    lib.amx_plan_create_freewater.argtypes = [i32, i32, i32, i32, vp, i32, vp, i32, vp, C.POINTER(vp)]
    lib.amx_plan_create_czb.argtypes = [i32, i32, i32, i32, vp, i32, vp, i32, vp, vp, vp, C.POINTER(vp)]
    lib.amx_plan_create_sandi.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.amx_plan_destroy.argtypes = [vp]
    lib.amx_plan_info.argtypes = [vp] + [C.POINTER(i32)] * 6
    lib.amx_fit.argtypes = [vp, C.POINTER(FitArgs), C.POINTER(i64)]
    lib.amx_lut_indices.argtypes = [vp, i32, vp, i64, vp]
    lib.amx_plan_last_timing.argtypes = [vp, C.POINTER(dbl), i32]
    lib.amx_plan_last_counters.argtypes = [vp, C.POINTER(i64), i32]
    lib.amx_preprocess.argtypes = [C.POINTER(PreArgs), C.POINTER(i64), C.POINTER(i32)]
    lib.amx_mean_b0.argtypes = [i32, i32, vp, i64, i32, vp, i32, vp, vp]
    lib.amx_dti_directions.argtypes = [i32, i32, vp, i32, i64, i32, vp, dbl, vp, vp]
    lib.amx_dti_directions_wls.argtypes = [i32, i32, vp, i32, i64, i32, vp, vp, dbl, vp, vp]
    lib.amx_scatter_maps.argtypes = [i32, i32, vp, i64, i32, vp, vp, i64, vp]
    lib.amx_resample_kernels.argtypes = [i32, i32, vp, i64, i32, vp, vp, i32, vp, i32, i32, vp, vp]
    lib.amx_volume_to_voxel_major.argtypes = [i32, i32, vp, i32, i64, i32, dbl, dbl, vp, vp]
    for name in EXPORTS:
        if name not in ("amx_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != AMX_OK:
        msg = load().amx_last_error().decode("utf-8", "replace")
        raise AmxError(rc, msg)
