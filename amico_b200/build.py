"""In-tree nvcc build of ``libamico_b200.so`` for sm_100a.

    python -m amico_b200.build [--force]

nvcc cross-compiles without a GPU, so this also runs in CI containers.  -fmad=false: the solver
arithmetic mirrors the reference's un-fused CPU arithmetic and uses explicit fma() where fusing is safe.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libamico_b200.so")
SOURCES = ["amx_api.cu", "amx_pre.cu"]
HEADERS = ["amx_err.h", "amx_warp.cuh", "amx_solvers.cuh", "amx_kernels.cuh", "amx_lean.cuh", "amx_slow.cuh", "amx_exact.cuh", "amx_small.cuh", os.path.join("..", "..", "include", "amico_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "2"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
