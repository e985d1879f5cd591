"""Build ``oracle/_ref``: the reference's own Cython hot-path glue, compiled from the sources where
they lie under /root/reference (nothing is copied into the repo; only binaries and our own stubs
land in the git-ignored ``oracle/_ref/``).

TEST INFRASTRUCTURE ONLY (see ``amico_oracle.c``).  What this gives:

* ``amico.models`` / ``amico.lut`` of daducci/AMICO cythonized UNMODIFIED, and the three pure-Python modules they import
  (``util``, ``scheme``, ``synthesis``) cythonized the same way -- binaries only, so that everything travels to the GPU
  box like any other built ``.so`` (a sourceless ``.pyc`` may be dropped by a snapshot filter): every line of chunking, dictionary assembly, NODDI stage logic, clamps, map formulas
  and fit errors is the reference's;
* the two solver entry points it cimports from the absent third-party ``spams-cython``
  (``cyspams.interfaces.nnls`` / ``.lasso``, amico/models.pyx:18) are bound to the restated
  solvers of ``amico_oracle.c`` through a shim ``cyspams`` package -- so this arm validates the
  oracle's *glue* bit-for-bit and serves as the "reference glue + restated solvers" CPU baseline,
  it does NOT pin the solvers themselves;
* ``dicelib.ui.ProgressBar`` and the three dipy symbols ``amico.lut`` imports are stubbed at import time by
  ``ref_runner._install_stubs`` (they are only exercised by kernel generation, which is outside the hot path).

Run:  python oracle/build_ref.py        (needs /root/reference, Cython, g++)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AMICO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

STUBS = {
    "amico/__init__.py": '"""Shim package root: the compiled reference modules live next to this file."""\n',
    "cyspams/__init__.py": '"""Shim for spams-cython: solver entry points bound to oracle/amico_oracle.c."""\n',
    "cyspams/interfaces.pxd": textwrap.dedent("""\
        # Signatures inferred from the reference's call sites (amico/models.pyx:615, 911, 926, 940, 1238, 1569).
        cdef extern from "cyspams_shim.h" nogil:
            void nnls(double* A, double* y, int m, int n, double* x, double& rnorm)
            void lasso(double* A, double* y, int m, int n, int p, double* x, double lambda1, double lambda2)
        """),
    "cyspams/cyspams_shim.h": textwrap.dedent("""\
        #pragma once
        extern "C" {
        int orc_nnls(const double*, const double*, int, int, double*, double*);
        int orc_lasso(const double*, const double*, int, int, int, double*, double, double);
        }
        static inline void nnls(double* A, double* y, int m, int n, double* x, double& rnorm) {
            orc_nnls(A, y, m, n, x, &rnorm);
        }
        static inline void lasso(double* A, double* y, int m, int n, int p, double* x, double l1, double l2) {
            orc_lasso(A, y, m, n, p, x, l1, l2);
        }
        """),
}

PY_MODULES = ["util.py", "synthesis.py", "scheme.py"]  # pure Python in the reference: cythonized to extension modules too
MODULES = ["lut.pyx", "models.pyx"]


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build():
    if not os.path.isdir(os.path.join(REF, "amico")):
        raise SystemExit(f"reference not found at {REF}: oracle/_ref can only be (re)built where it is mounted")
    for rel, text in STUBS.items():
        p = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
    bdir = os.path.join(OUT, "build")
    os.makedirs(bdir, exist_ok=True)
    inc = sysconfig.get_paths()["include"]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    obj = os.path.join(bdir, "amico_oracle.o")
    run(["gcc", "-O2", "-fPIC", "-std=c99", "-c", os.path.join(HERE, "amico_oracle.c"), "-o", obj])
    for stale in os.listdir(os.path.join(OUT, "amico")):
        if stale.endswith(".pyc"):
            os.remove(os.path.join(OUT, "amico", stale))
    for mod in PY_MODULES + MODULES:
        name = mod.split(".")[0]
        cpp = os.path.join(bdir, name + ".cpp")
        src = os.path.join(REF, "amico", mod)
        if mod.endswith(".py"):
            # CPython accepts the reference's util.py (tabs in some functions, spaces in others), Cython does not: compile a
            # scratch copy with the tabs expanded (whitespace only; lives in the build directory that is removed below)
            with open(src) as f:
                text = f.read().expandtabs(4)
            src = os.path.join(bdir, mod)
            with open(src, "w") as f:
                f.write(text)
        # --module-name: the .py files would otherwise be compiled as top-level modules
        run([sys.executable, "-m", "cython", "--cplus", "-3", "-I", OUT, "-I", REF, "--module-name", "amico." + name, src, "-o", cpp])
        so = os.path.join(OUT, "amico", name + ext)
        # the reference builds with -std=c++14 -Ofast (setup.py:40); -O3 keeps IEEE semantics so
        # that this arm is comparable bit-for-bit with the plain-C oracle
        cmd = ["g++", "-O3", "-std=c++14", "-fPIC", "-shared", "-w", "-I", inc, "-I", os.path.join(OUT, "cyspams"),
               cpp, "-o", so]
        if name == "models":
            cmd += [obj, "-lpthread", "-lm"]
        run(cmd)
    import shutil
    shutil.rmtree(bdir, ignore_errors=True)  # generated C++ is large and need not travel to the GPU box
    print("built", OUT)


if __name__ == "__main__":
    build()
