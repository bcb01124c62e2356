"""ctypes binding of the host-side proposal generator (plum_b200/host/mc_propose.h through the pmc_* entry
points of libplum_mcbench.so).  Tests and tools use it to walk a std::mt19937 in the reference's draw order
(src/simulation/simulation.cc:216-355) and to build trial coordinates with the host version of k_propose."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._abi import PgMoveDesc, c_double_p, c_int32_p, dptr, iptr

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplum_mcbench.so")
KIND_NONE, KIND_GC, KIND_UNSUPPORTED = -1, -2, -3
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.pmc_create.restype = C.c_void_p
        L.pmc_create.argtypes = [C.c_int, c_int32_p, C.c_int, C.c_double, c_double_p, C.c_double, C.c_int, C.c_int, C.c_uint]
        L.pmc_destroy.argtypes = [C.c_void_p]
        L.pmc_next.argtypes = [C.c_void_p, C.POINTER(PgMoveDesc), c_double_p, C.c_int, C.POINTER(C.c_int)]
        L.pmc_no_accept_draw.argtypes = [C.c_void_p]
        L.pmc_no_accept_draw.restype = None
        L.pmc_peek_raw.argtypes = [C.c_void_p]
        L.pmc_peek_raw.restype = C.c_uint
        L.pmc_apply.argtypes = [C.POINTER(PgMoveDesc), c_double_p, C.c_int, c_double_p, c_double_p]
        L.pmc_apply.restype = None
        _LIB = L
    return _LIB


def bond_settings(r):
    """(bond_len, vary_bond) as Simulation::TranslationalMove picks them (simulation.cc:288-296)."""
    bond_len, vary = 0.0, False
    if r.use_rigid:
        bond_len = r.rigid_bond
    if r.use_bond:
        bond_len, vary = r.bond_r0, True
    return bond_len, vary


class Generator:
    def __init__(self, mol_len, phantom, move_size, move_prob, bond_len, vary_bond, gc_freq, seed, max_len=None):
        self.L = lib()
        ml = np.ascontiguousarray(mol_len, dtype=np.int32)
        pr = np.ascontiguousarray(move_prob, dtype=np.float64)
        self.h = self.L.pmc_create(len(ml), iptr(ml), int(phantom), float(move_size), dptr(pr), float(bond_len),
                                   int(bool(vary_bond)), int(gc_freq), int(seed))
        self.cap = int(max_len or max(int(ml.max()) if len(ml) else 1, 1))
        self._rv = np.zeros((self.cap, 4))

    @classmethod
    def for_run(cls, r, mol_first, seed):
        bl, vary = bond_settings(r)
        return cls(np.diff(mol_first), r.phantom, r.move_size, r.move_prob, bl, vary, r.gc_freq if r.use_gc else 0, seed)

    def close(self):
        if self.h:
            self.L.pmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def next(self):
        """One step: (kind, desc, rvec rows).  kind < 0: KIND_NONE / KIND_GC / KIND_UNSUPPORTED."""
        d = PgMoveDesc()
        n = C.c_int(0)
        kind = self.L.pmc_next(self.h, C.byref(d), dptr(self._rv), self.cap, C.byref(n))
        if kind == -4:
            raise RuntimeError("pivot rows exceed the generator's row buffer")
        return kind, d, self._rv[: n.value].copy()

    def no_accept_draw(self):
        self.L.pmc_no_accept_draw(self.h)

    def peek_raw(self) -> int:
        return int(self.L.pmc_peek_raw(self.h))


def apply(desc: PgMoveDesc, rvec: np.ndarray, cur: np.ndarray) -> np.ndarray:
    cur = np.ascontiguousarray(cur, dtype=np.float64)
    out = np.zeros_like(cur)
    rv = np.ascontiguousarray(rvec, dtype=np.float64).reshape(-1, 4)
    d = PgMoveDesc()
    C.memmove(C.byref(d), C.byref(desc), C.sizeof(PgMoveDesc))
    d.rv_offset = 0
    lib().pmc_apply(C.byref(d), dptr(rv) if rv.size else None, int(cur.shape[0]), dptr(cur), dptr(out))
    return out
