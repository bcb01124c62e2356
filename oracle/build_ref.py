#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — builds the reference's own CPU implementation.

Compiles /root/reference/src (nuwapi/Plum, unmodified arithmetic) into
``oracle/_ref/plum_ref`` so the restatement in ``oracle/plum_oracle.c`` and the
CUDA path can be pinned against the real thing.  Nothing from the reference is
copied into the repository: the sources are copied to a throw-away directory
under /tmp, observation hooks are applied to that copy, and only the binary lands in
``oracle/_ref/`` (git-ignored).

Differences from the reference's own ``src/Makefile:1-25``:
  * ``-lfftw3`` dropped (no fftw symbol is referenced anywhere in src/);
  * ``-I oracle/eigen_standin`` instead of a real Eigen (see that header);
  * hook 1 (seed):  ``PLUM_SEED`` overrides the wall-clock seed at
    ``src/simulation/simulation.cc:134``;
  * hook 2 (trace): when ``PLUM_TRACE`` names a file, one line per MC step is
    appended with the move, dE (hex float), the accept decision and the four
    running energy totals, and one per volume-perturbation pressure sample (the
    accumulators of src/force_field/pressure.cc:187-387, which the reference never
    prints for bulk systems).  The hooks only *observe*; no arithmetic changes.

Trace format: see plum_b200/host/driver_hooks.py.
"""
import os
import shutil
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plum_b200.host import driver_hooks  # noqa: E402  (the hooks are shared with bin/plum_gpu)

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLUM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"build_ref: {REF}/src not present — keeping any prebuilt oracle/_ref/plum_ref")
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="plum_ref_build_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(os.path.join(REF, "src"), src)
        driver_hooks.apply_sim_hooks(src)
        driver_hooks.apply_cbmc_hooks(src)
        driver_hooks.apply_pressure_hook(src)
        files = [os.path.join(src, "main.cc")]
        for d in ("simulation", "molecules", "utilities", "force_field"):
            files += sorted(os.path.join(src, d, f) for f in os.listdir(os.path.join(src, d)) if f.endswith(".cc"))
        cmd = ["g++", "-std=c++11", "-O3", "-w", "-ffp-contract=off", "-I", os.environ.get("EIGEN_INCLUDE", os.path.join(os.path.dirname(HERE), "plum_b200", "host", "eigen_standin")),
               "-o", os.path.join(OUT, "plum_ref")] + files + ["-lm"]
        subprocess.check_call(cmd)
        print("build_ref: built", os.path.join(OUT, "plum_ref"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
