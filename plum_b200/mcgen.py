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
    if d.kind == 2:      # PIVOT: the rows handed in start at 0 (CRANKSHAFT keeps its last-bead index there)
        d.rv_offset = 0
    lib().pmc_apply(C.byref(d), dptr(rv) if rv.size else None, int(cur.shape[0]), dptr(cur), dptr(out))
    return out


def _bench_lib():
    L = lib()
    if not getattr(L, "_pb_ready", False):
        dp, ip, bp = c_double_p, c_int32_p, C.POINTER(C.c_uint8)
        L.pb_create.restype = C.c_void_p
        L.pb_create.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, C.c_double, C.c_double, C.c_double, dp, C.c_int, C.c_uint]
        L.pb_destroy.argtypes = [C.c_void_p]
        L.pb_set_records.argtypes = [C.c_void_p, ip, ip, dp, dp, bp, dp, bp, C.c_int]
        L.pb_set_records.restype = None
        L.pb_set_vary_bond.argtypes = [C.c_void_p, C.c_int]
        L.pb_set_vary_bond.restype = None
        L.pb_positions.argtypes = [C.c_void_p, dp]
        L.pb_positions.restype = None
        L.pb_stats.argtypes = [C.c_void_p, ip, dp, dp, ip]
        L.pb_stats.restype = None
        L.pb_run_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp]
        L.pb_run_mc.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, dp]
        L.pb_run_chain.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.pb_chain_seed.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
        L.pb_chain_resident.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L._pb_ready = True
    return L


class NativeChain:
    """One Markov chain driven by the native caller (plum_b200/host/mc_bench.cc) over the C ABI: either move by
    move (pg_delta_e / pg_commit, trial coordinates built on the host) or in batches with device-side proposals
    (pg_mc_*).  Both walk the same std::mt19937 in the reference's draw order, so for one seed they must produce
    the same chain."""

    def __init__(self, eng, r, sysm, seed, record_moves=0, record_trials=False):
        self.L = _bench_lib()
        self.eng = eng
        self.n = sysm.n
        mf = np.ascontiguousarray(sysm.mol_first, dtype=np.int32)
        xyz = np.ascontiguousarray(sysm.xyz, dtype=np.float64)
        box = np.ascontiguousarray(sysm.box, dtype=np.float64)
        prob = np.ascontiguousarray(r.move_prob, dtype=np.float64)
        bl, vary = bond_settings(r)
        self.ctx = self.L.pb_create(eng.h, sysm.n_mol, iptr(mf), dptr(xyz), dptr(box), float(r.beta), float(r.move_size),
                                    float(bl), dptr(prob), int(r.phantom), int(seed))
        self.L.pb_set_vary_bond(self.ctx, int(vary))
        self.max_len = int(np.diff(mf).max())
        self.cap = record_moves
        if record_moves:
            self.rec_mol = np.zeros(record_moves, dtype=np.int32)
            self.rec_off = np.zeros(record_moves, dtype=np.int32)
            self.rec_u = np.zeros(record_moves)
            self.rec_dE = np.zeros(record_moves)
            self.rec_acc = np.zeros(record_moves, dtype=np.uint8)
            nb = record_moves * self.max_len if record_trials else 0
            self.rec_trial = np.zeros((max(nb, 1), 3))
            self.rec_moved = np.zeros(max(nb, 1), dtype=np.uint8)
            bp = C.POINTER(C.c_uint8)
            self.L.pb_set_records(self.ctx, iptr(self.rec_mol), iptr(self.rec_off), dptr(self.rec_u), dptr(self.rec_dE),
                                  self.rec_acc.ctypes.data_as(bp), dptr(self.rec_trial),
                                  self.rec_moved.ctypes.data_as(bp), nb)

    def close(self):
        if self.ctx:
            self.L.pb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, fn, *args):
        arr = (C.c_void_p * 1)(self.ctx)
        wall = C.c_double(0.0)
        rc = fn(arr, 1, *args, C.byref(wall))
        if rc != 0:
            raise RuntimeError(f"native chain failed ({rc}): {self.eng.L.pg_last_error(self.eng.h).decode()}")
        return wall.value

    def run_per_move(self, n_moves):
        """n_moves Metropolis steps, one pg_delta_e + pg_commit round trip each.  Records restart at 0."""
        return self._run(self.L.pb_run_multi, int(n_moves))

    def run_batched(self, n_moves, batch_moves=1024):
        """The same steps with device-side proposals, `batch_moves` per upload."""
        return self._run(self.L.pb_run_mc, int(n_moves), int(batch_moves))

    def run_chain(self, n_moves, batch=1024, cluster=1, keep_trials=False):
        """The same steps device-resident (pg_chain_*): the generator state crosses the boundary, nothing else."""
        return self._run(self.L.pb_run_chain, int(n_moves), int(batch), int(cluster), int(bool(keep_trials)), 1)

    def positions(self):
        out = np.zeros((self.n, 3))
        self.L.pb_positions(self.ctx, dptr(out))
        return out


def run_chain_multi(chains, n_moves, batch=1024, cluster=1):
    """Several NativeChain replicas (one engine each, same device) through pg_chain_run_multi: one launch per batch
    carries all of them, one host thread drives everything."""
    L = _bench_lib()
    arr = (C.c_void_p * len(chains))(*[c.ctx for c in chains])
    wall = C.c_double(0.0)
    rc = L.pb_run_chain(arr, len(chains), int(n_moves), int(batch), int(cluster), 0, 1, C.byref(wall))
    if rc != 0:
        msgs = [c.eng.L.pg_last_error(c.eng.h).decode() for c in chains]
        raise RuntimeError(f"pb_run_chain failed ({rc}): {msgs}")
    return wall.value
