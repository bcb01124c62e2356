"""TEST INFRASTRUCTURE ONLY — ctypes wrapper over oracle/libplum_oracle.so.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
The interface mirrors plum_b200.engine.Engine so a test can drive both with the
same calls and compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from plum_b200._abi import (ParamBlock, PgDelta, PgEwaldInfo, PgParams, PgTotals, bptr, c_double_p, c_int32_p,
                            c_uint8_p, dptr, iptr)

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libplum_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libplum_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "plum_oracle.c")):
            build()
        L = C.CDLL(path)
        L.po_create.restype = C.c_void_p
        L.po_create.argtypes = [C.POINTER(PgParams)]
        L.po_destroy.argtypes = [C.c_void_p]
        L.po_set_repl_mode.argtypes = [C.c_void_p, C.c_int]
        L.po_set_timing_new_only.argtypes = [C.c_void_p, C.c_int]
        L.po_get_ewald_info.argtypes = [C.c_void_p, C.POINTER(PgEwaldInfo)]
        L.po_upload_system.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_int32_p, C.c_int, c_int32_p]
        L.po_num_beads.argtypes = [C.c_void_p]
        L.po_download_positions.argtypes = [C.c_void_p, c_double_p]
        L.po_compute_totals.argtypes = [C.c_void_p, C.POINTER(PgTotals)]
        L.po_init_energy.argtypes = [C.c_void_p, C.POINTER(PgTotals)]
        L.po_get_totals.argtypes = [C.c_void_p, C.POINTER(PgTotals)]
        L.po_delta_e.argtypes = [C.c_void_p, C.c_int, c_double_p, c_uint8_p, C.POINTER(PgDelta)]
        L.po_commit.argtypes = [C.c_void_p, C.c_int]
        L.po_mz_current.restype = C.c_double
        L.po_mz_current.argtypes = [C.c_void_p]
        L.po_vol_scaling.restype = C.c_int
        L.po_vol_scaling.argtypes = [C.c_void_p, C.c_int, C.c_double, c_double_p]
        L.po_wall_force.restype = C.c_int
        L.po_wall_force.argtypes = [C.c_void_p, C.c_int, c_double_p]
        L.po_beads_energy.restype = C.c_double
        L.po_beads_energy.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_double, c_double_p, C.c_int, C.c_double,
                                      C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p, C.c_int, C.c_int,
                                      c_double_p, c_double_p]
        L.po_insert_molecules.argtypes = [C.c_void_p, C.c_int, c_int32_p, c_double_p, c_double_p, c_int32_p,
                                          C.POINTER(PgTotals)]
        L.po_delete_molecules.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PgTotals)]
        L.po_k_half_list.argtypes = [C.c_void_p, c_int32_p, c_double_p, C.c_int]
        L.po_sk_half.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int]
        for f in ("po_pair_energy",):
            getattr(L, f).restype = C.c_double
        L.po_pair_energy.argtypes = [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int]
        L.po_pair_real.restype = C.c_double
        L.po_pair_real.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p, C.c_double]
        L.po_pair_repl.restype = C.c_double
        L.po_pair_repl.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p, C.c_double]
        L.po_wall_energy.restype = C.c_double
        L.po_wall_energy.argtypes = [C.c_void_p, c_double_p, C.c_int]
        _LIB = L
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, params: dict, repl_mode: int = 0):
        self.pb = ParamBlock(params)
        self.L = lib()
        self.h = self.L.po_create(C.byref(self.pb.struct))
        self.L.po_set_repl_mode(self.h, repl_mode)
        self.mol_first = np.zeros(1, dtype=np.int32)

    def close(self):
        if self.h:
            self.L.po_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_repl_mode(self, mode: int):
        self.L.po_set_repl_mode(self.h, mode)

    def set_timing_new_only(self, flag: bool):
        """Bench only: evaluate new-configuration energies only (the reference's per-move work)."""
        self.L.po_set_timing_new_only(self.h, int(bool(flag)))

    def ewald_info(self) -> PgEwaldInfo:
        o = PgEwaldInfo()
        self.L.po_get_ewald_info(self.h, C.byref(o))
        return o

    def upload(self, xyz, q, type_ids, mol_first):
        xyz = _f64(xyz).reshape(-1, 3)
        q = _f64(q)
        t = np.ascontiguousarray(type_ids, dtype=np.int32)
        mf = np.ascontiguousarray(mol_first, dtype=np.int32)
        self.mol_first = mf.copy()
        self.L.po_upload_system(self.h, xyz.shape[0], dptr(xyz), dptr(q), iptr(t), mf.shape[0] - 1, iptr(mf))

    @property
    def n(self):
        return self.L.po_num_beads(self.h)

    def positions(self):
        out = np.zeros((self.n, 3), dtype=np.float64)
        if self.n:
            self.L.po_download_positions(self.h, dptr(out))
        return out

    def compute_totals(self) -> dict:
        t = PgTotals()
        self.L.po_compute_totals(self.h, C.byref(t))
        return t.as_dict()

    def init_energy(self) -> dict:
        t = PgTotals()
        self.L.po_init_energy(self.h, C.byref(t))
        return t.as_dict()

    def totals(self) -> dict:
        t = PgTotals()
        self.L.po_get_totals(self.h, C.byref(t))
        return t.as_dict()

    def delta_e(self, mol: int, trial_xyz, moved) -> dict:
        xyz = _f64(trial_xyz).reshape(-1, 3)
        mv = np.ascontiguousarray(moved, dtype=np.uint8)
        d = PgDelta()
        self.L.po_delta_e(self.h, int(mol), dptr(xyz), bptr(mv), C.byref(d))
        return d.as_dict()

    def commit(self, accept: bool):
        rc = self.L.po_commit(self.h, int(bool(accept)))
        assert rc == 0, rc

    def vol_scaling_sample(self, phantom: int, dz: float = 1e-5) -> dict:
        """One sample of CalcPressureVolScalingHSELSlit's dU terms (pressure.cc:187-338), pairwise like the reference."""
        out = np.zeros(36)
        rc = self.L.po_vol_scaling(self.h, int(phantom), float(dz), dptr(out))
        if rc != 0:
            raise RuntimeError("po_vol_scaling failed")
        return {"el": out[:16].copy(), "hs": out[16:32].copy(), "bond": float(out[32]), "dipole": float(out[33]),
                "dU": float(out[34]), "n_free": int(out[35])}

    def wall_force(self, phantom: int) -> np.ndarray:
        """One sample of CalcPressureForceLJELSlit (pressure.cc:404-469): the six force sums
        {LJ ion, LJ polymer, LJ wall-wall, EL ion, EL polymer, EL wall-wall}."""
        out = np.zeros(6)
        rc = self.L.po_wall_force(self.h, int(phantom), dptr(out))
        if rc:
            raise RuntimeError("po_wall_force failed")
        return out

    def beads_energy(self, b1, t1, q1, b2, t2, q2, use_bead2, chain_xyz, chain_q, chain_type, current_len,
                     skip_first=-1, skip_last=-1):
        b1 = _f64(b1)
        b2 = _f64(b2)
        cx = _f64(chain_xyz).reshape(-1, 3) if current_len else np.zeros((1, 3))
        cq = _f64(chain_q) if current_len else np.zeros(1)
        ct = np.ascontiguousarray(chain_type, dtype=np.int32) if current_len else np.zeros(1, dtype=np.int32)
        pe = C.c_double()
        ee = C.c_double()
        e = self.L.po_beads_energy(self.h, dptr(b1), int(t1), float(q1), dptr(b2), int(t2), float(q2), int(use_bead2),
                                   int(current_len), dptr(cx), dptr(cq), iptr(ct), int(skip_first), int(skip_last),
                                   C.cast(C.byref(pe), c_double_p), C.cast(C.byref(ee), c_double_p))
        return e, pe.value, ee.value

    def insert_molecules(self, mol_len, xyz, q, type_ids) -> dict:
        ml = np.ascontiguousarray(mol_len, dtype=np.int32)
        xyz = _f64(xyz).reshape(-1, 3)
        q = _f64(q)
        t = np.ascontiguousarray(type_ids, dtype=np.int32)
        a = PgTotals()
        self.L.po_insert_molecules(self.h, ml.shape[0], iptr(ml), dptr(xyz), dptr(q), iptr(t), C.byref(a))
        last = int(self.mol_first[-1])
        self.mol_first = np.concatenate([self.mol_first, last + np.cumsum(ml)]).astype(np.int32)
        return a.as_dict()

    def delete_molecules(self, mf: int, ml: int) -> dict:
        r = PgTotals()
        self.L.po_delete_molecules(self.h, int(mf), int(ml), C.byref(r))
        first = self.mol_first
        nrem = int(first[ml + 1] - first[mf])
        self.mol_first = np.concatenate([first[:mf + 1], first[ml + 2:] - nrem]).astype(np.int32)
        return r.as_dict()

    def k_half_list(self):
        n = self.L.po_k_half_list(self.h, None, None, 0)
        l = np.zeros((max(n, 1), 3), dtype=np.int32)
        e = np.zeros(max(n, 1), dtype=np.float64)
        self.L.po_k_half_list(self.h, iptr(l), dptr(e), n)
        return l[:n], e[:n]

    def sk_half(self, xyz=None):
        pos = _f64(self.positions() if xyz is None else xyz).reshape(-1, 3)
        n = self.L.po_k_half_list(self.h, None, None, 0)
        out = np.zeros((max(n, 1), 2), dtype=np.float64)
        self.L.po_sk_half(self.h, dptr(pos), dptr(out), n)
        return out[:n]

    # primitives
    def pair_energy(self, a, ta, b, tb):
        return self.L.po_pair_energy(self.h, dptr(_f64(a)), ta, dptr(_f64(b)), tb)

    def pair_real(self, a, qa, b, qb):
        return self.L.po_pair_real(self.h, dptr(_f64(a)), qa, dptr(_f64(b)), qb)

    def pair_repl(self, a, qa, b, qb):
        return self.L.po_pair_repl(self.h, dptr(_f64(a)), qa, dptr(_f64(b)), qb)

    def wall_energy(self, p, t):
        return self.L.po_wall_energy(self.h, dptr(_f64(p)), t)
