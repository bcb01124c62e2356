#!/usr/bin/env python3
"""Builds bin/plum_gpu: Plum's own MC driver, unchanged, on top of the B200 façade.

  driver    main.cc, simulation/, molecules/, utilities/  — compiled from the reference tree
            (PLUM_REFERENCE, default /root/reference) where it lies; copied to a throw-away
            directory only so that `#include "../force_field/force_field.h"` resolves to the
            façade header and the seed/trace hooks (driver_hooks.py) can be applied;
  façade    plum_b200/host/force_field.{h,cc}  (class ForceField, same public interface);
  kernels   plum_b200/libplum_b200.so through include/plum_b200.h.

Nothing from the reference is written into the repository; only the binary lands in bin/
(git-ignored).  Without the reference tree (e.g. on the GPU box) a prebuilt bin/plum_gpu is kept.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PLUM_REFERENCE", "/root/reference")
PROFILE = bool(os.environ.get("PLUM_HOST_GPROF"))   # -pg build for gprof (bin/plum_gpu_prof)
OUT = os.path.join(REPO, "bin", "plum_gpu_prof" if PROFILE else "plum_gpu")
sys.path.insert(0, REPO)
from plum_b200.host import driver_hooks  # noqa: E402


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in ("force_field.h", "force_field.cc", "driver_hooks.py", "build_host.py", "mc_propose.h", "mt_state.h")]
    deps.append(os.path.join(REPO, "plum_b200", "csrc", "pg_propose_math.h"))
    deps.append(os.path.join(REPO, "include", "plum_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def main():
    force = "--force" in sys.argv
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"build_host: {REF}/src not present — keeping any prebuilt bin/plum_gpu")
        return 0
    if not force and not stale():
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="plum_gpu_build_")
    try:
        src = os.path.join(tmp, "src")
        os.makedirs(src)
        shutil.copy(os.path.join(REF, "src", "main.cc"), src)
        for d in ("simulation", "molecules", "utilities"):
            shutil.copytree(os.path.join(REF, "src", d), os.path.join(src, d))
        os.makedirs(os.path.join(src, "force_field"))
        for f in ("force_field.h", "force_field.cc"):
            shutil.copy(os.path.join(HERE, f), os.path.join(src, "force_field", f))
        driver_hooks.apply_sim_hooks(src)
        driver_hooks.apply_batch_hook(src)
        files = [os.path.join(src, "main.cc"), os.path.join(src, "force_field", "force_field.cc")]
        for d in ("simulation", "molecules", "utilities"):
            files += sorted(os.path.join(src, d, f) for f in os.listdir(os.path.join(src, d)) if f.endswith(".cc"))
        eigen = os.environ.get("EIGEN_INCLUDE", os.path.join(HERE, "eigen_standin"))
        libdir = os.path.join(REPO, "plum_b200")
        cmd = ["g++", "-std=c++11", "-O3", "-w"] + (["-pg", "-fno-inline-functions"] if PROFILE else []) + ["-ffp-contract=off", "-I", eigen, "-I", os.path.join(REPO, "include"), "-I", HERE, "-I", os.path.join(REPO, "plum_b200", "csrc"), "-o", OUT] + files + \
              ["-L", libdir, "-lplum_b200", "-Wl,-rpath,$ORIGIN/../plum_b200", "-lm"]
        subprocess.check_call(cmd)
        print("build_host: built", OUT)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
