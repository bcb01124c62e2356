"""plum_b200 — B200-native energy engine for Plum's per-trial-move energy path.

Layout:
  csrc/     sm_100a CUDA kernels + the C ABI (include/plum_b200.h) -> libplum_b200.so
  host/     C++ façade with the reference's ForceField public interface
  engine.py ctypes binding of the C ABI (tests, bench)
  runin.py  readers for Plum's own input formats
"""
__all__ = ["engine", "runin"]
