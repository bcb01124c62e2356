"""ctypes mirror of include/plum_b200.h (struct layouts only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)

VERY_LARGE_ENERGY = 1.0e8


class PgParams(C.Structure):
    _fields_ = [
        ("box", C.c_double * 3), ("npbc", C.c_int32), ("n_types", C.c_int32), ("beta", C.c_double),
        ("pair_kind", C.c_int32), ("_pad0", C.c_int32), ("lj_cutoff", C.c_double),
        ("lj_sigma", c_double_p), ("lj_epsilon", c_double_p), ("hs_radius", c_double_p),
        ("use_ewald", C.c_int32), ("dipole_correction", C.c_int32), ("lB", C.c_double), ("alpha", C.c_double),
        ("bond_kind", C.c_int32), ("_pad1", C.c_int32), ("bond_k", C.c_double), ("bond_r0", C.c_double),
        ("ext_kind", C.c_int32), ("_pad2", C.c_int32), ("wall_cut", C.c_double),
        ("wall_sigma", c_double_p), ("wall_epsilon", c_double_p), ("graft_kind", c_int32_p),
        ("well_width", C.c_double), ("well_depth", C.c_double),
    ]


class PgEwaldInfo(C.Structure):
    _fields_ = [
        ("ewald_box", C.c_double * 3), ("box_vol", C.c_double), ("real_cutoff", C.c_double),
        ("real_cell", C.c_int32 * 3), ("n_k", C.c_int32), ("repl_cutoff", C.c_double),
        ("repl_cell", C.c_int32 * 3), ("n_k_half", C.c_int32),
    ]


class PgDelta(C.Structure):
    _fields_ = [
        ("dE", C.c_double), ("pair", C.c_double), ("ext", C.c_double), ("ewald", C.c_double),
        ("bond", C.c_double), ("real", C.c_double), ("recip", C.c_double), ("mz_current", C.c_double),
        ("stage", C.c_int32), ("n_overlap", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PgTotals(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("pair", "ewald", "bond", "ext", "real", "recip", "self", "dipole")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PgProposal(C.Structure):
    _fields_ = [("mol", C.c_int32), ("xyz_offset", C.c_int32), ("u", C.c_double)]


class PgMoveDesc(C.Structure):
    _fields_ = [("mol", C.c_int32), ("kind", C.c_int32), ("i0", C.c_int32), ("rv_offset", C.c_int32),
                ("s", C.c_double), ("v", C.c_double * 3), ("vlen", C.c_double), ("u", C.c_double)]


class PgChainConfig(C.Structure):
    _fields_ = [("phantom", C.c_int32), ("gc_freq", C.c_int32), ("vary_bond", C.c_int32), ("cluster_ctas", C.c_int32),
                ("keep_trials", C.c_int32), ("pivot_mode", C.c_int32), ("move_size", C.c_double), ("bond_len", C.c_double),
                ("move_prob", C.c_double * 5)]


class PgChainStep(C.Structure):
    _fields_ = [("dE", C.c_double), ("mol", C.c_int32), ("kind", C.c_int8), ("accept", C.c_uint8), ("stage", C.c_uint8),
                ("_pad", C.c_uint8)]


class PgVolSample(C.Structure):
    _fields_ = [("el", C.c_double * 16), ("hs", C.c_double * 16), ("bond", C.c_double), ("dipole", C.c_double),
                ("dU", C.c_double), ("n_free", C.c_int32), ("_pad", C.c_int32)]


class PgTrialSet(C.Structure):
    _fields_ = [
        ("n_trials", C.c_int32), ("use_bead2", C.c_int32), ("type1", C.c_int32), ("type2", C.c_int32),
        ("q1", C.c_double), ("q2", C.c_double), ("current_len", C.c_int32),
        ("skip_mol_first", C.c_int32), ("skip_mol_last", C.c_int32), ("_pad", C.c_int32),
    ]


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


def bptr(a: np.ndarray):
    return a.ctypes.data_as(c_uint8_p)


class ParamBlock:
    """Owns the numpy tables a PgParams points into."""

    def __init__(self, d: dict):
        self._keep = {}
        p = PgParams()
        for i in range(3):
            p.box[i] = d["box"][i]
        for k in ("npbc", "n_types", "beta", "pair_kind", "lj_cutoff", "use_ewald", "dipole_correction", "lB", "alpha",
                  "bond_kind", "bond_k", "bond_r0", "ext_kind", "wall_cut", "well_width", "well_depth"):
            setattr(p, k, d[k])
        for k in ("lj_sigma", "lj_epsilon", "hs_radius", "wall_sigma", "wall_epsilon"):
            a = np.ascontiguousarray(d[k], dtype=np.float64)
            self._keep[k] = a
            setattr(p, k, dptr(a))
        g = np.ascontiguousarray(d["graft_kind"], dtype=np.int32)
        self._keep["graft_kind"] = g
        p.graft_kind = iptr(g)
        self.struct = p
        self.dict = d
