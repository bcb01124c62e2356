"""ctypes binding of libplum_b200.so (the C ABI in include/plum_b200.h).

This is the Python harness over the boundary — used by tests and bench.py.  The
per-move product path is C++ (plum_b200/host/force_field.cc) over the same ABI.
There is no fallback: if the shared library or a CUDA device is missing, every
entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._abi import (ParamBlock, PgChainConfig, PgChainStep, PgDelta, PgEwaldInfo, PgMoveDesc, PgParams, PgProposal, PgTotals, PgTrialSet, PgVolSample, bptr, c_double_p,
                   c_int32_p, c_uint8_p, dptr, iptr)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplum_b200.so")

# Every symbol include/plum_b200.h declares (tests check the library exports them all).
ABI_SYMBOLS = [
    "pg_create", "pg_destroy", "pg_last_error", "pg_abi_version", "pg_get_ewald_info", "pg_upload_system",
    "pg_init_energy", "pg_recompute_totals", "pg_get_totals", "pg_download_positions", "pg_num_beads", "pg_delta_e",
    "pg_delta_e_begin", "pg_delta_e_poll", "pg_commit", "pg_replay_upload", "pg_replay_run", "pg_replay_time_delta", "pg_replay_prepare", "pg_mc_upload", "pg_mc_run", "pg_mc_begin", "pg_mc_end", "pg_mc_trial_xyz", "pg_trial_energies", "pg_insert_molecules",
    "pg_delete_molecules", "pg_wall_force", "pg_vol_scaling_sample", "pg_sk_compute_slice", "pg_sk_set", "pg_sk_energy", "pg_sk_download", "pg_launch_count",
    "pg_stream", "pg_measure_fp64_peak",
    "pg_chain_configure", "pg_chain_set_rng", "pg_chain_get_rng", "pg_chain_run", "pg_chain_begin", "pg_chain_end",
    "pg_chain_run_multi", "pg_chain_steps", "pg_chain_trial_xyz", "pg_chain_check", "pg_chain_counters",
    "pg_chain_run_multi_io", "pg_measure_l2_peak",
    "pg_recompute_sk", "pg_recompute_sk_begin", "pg_recompute_sk_end", "pg_sk_export", "pg_sk_attach", "pg_sk_attach_local", "pg_sk_detach",
]

_LIB = None


class EngineError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` — there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.pg_create.argtypes = [C.POINTER(PgParams), C.c_int, C.c_int, C.POINTER(vp)]
        L.pg_destroy.argtypes = [vp]
        L.pg_last_error.restype = C.c_char_p
        L.pg_last_error.argtypes = [vp]
        L.pg_get_ewald_info.argtypes = [vp, C.POINTER(PgEwaldInfo)]
        L.pg_upload_system.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_int32_p, C.c_int, c_int32_p]
        L.pg_init_energy.argtypes = [vp, C.POINTER(PgTotals)]
        L.pg_recompute_totals.argtypes = [vp, C.POINTER(PgTotals)]
        L.pg_get_totals.argtypes = [vp, C.POINTER(PgTotals)]
        L.pg_download_positions.argtypes = [vp, c_double_p]
        L.pg_num_beads.argtypes = [vp]
        L.pg_delta_e.argtypes = [vp, C.c_int, c_double_p, c_uint8_p, C.POINTER(PgDelta)]
        L.pg_wall_force.argtypes = [vp, C.c_int, c_double_p]
        L.pg_vol_scaling_sample.argtypes = [vp, C.c_int, C.c_double, C.POINTER(PgVolSample)]
        L.pg_delta_e_begin.argtypes = [vp, C.c_int, c_double_p, c_uint8_p]
        L.pg_delta_e_poll.argtypes = [vp, C.POINTER(PgDelta)]
        L.pg_commit.argtypes = [vp, C.c_int]
        L.pg_replay_upload.argtypes = [vp, C.c_int, C.POINTER(PgProposal), C.c_int, c_double_p, c_uint8_p]
        L.pg_replay_run.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_uint8_p, C.POINTER(C.c_float)]
        L.pg_replay_time_delta.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.pg_replay_prepare.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.pg_mc_upload.argtypes = [vp, C.c_int, C.POINTER(PgMoveDesc), C.c_int, c_double_p]
        L.pg_mc_run.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_uint8_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.pg_mc_begin.argtypes = [vp, C.c_int, C.c_int]
        L.pg_mc_end.argtypes = [vp, c_double_p, c_uint8_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.pg_mc_trial_xyz.argtypes = [vp, C.c_int, c_double_p]
        L.pg_trial_energies.argtypes = [vp, C.POINTER(PgTrialSet), c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_int32_p, c_double_p, c_double_p, c_double_p]
        L.pg_insert_molecules.argtypes = [vp, C.c_int, c_int32_p, c_double_p, c_double_p, c_int32_p,
                                          C.POINTER(PgTotals)]
        L.pg_delete_molecules.argtypes = [vp, C.c_int, C.c_int, C.POINTER(PgTotals)]
        L.pg_sk_compute_slice.argtypes = [vp, C.c_int, C.c_int, C.c_void_p]
        L.pg_sk_set.argtypes = [vp, C.c_void_p]
        L.pg_sk_energy.argtypes = [vp, C.c_int, C.c_int, c_double_p]
        L.pg_sk_download.argtypes = [vp, c_double_p]
        L.pg_launch_count.restype = C.c_uint64
        L.pg_launch_count.argtypes = [vp]
        L.pg_stream.restype = C.c_void_p
        L.pg_stream.argtypes = [vp]
        L.pg_measure_fp64_peak.argtypes = [vp, c_double_p]
        u32p = C.POINTER(C.c_uint32)
        L.pg_chain_configure.argtypes = [vp, C.POINTER(PgChainConfig)]
        L.pg_chain_set_rng.argtypes = [vp, u32p, C.c_int]
        L.pg_chain_get_rng.argtypes = [vp, u32p, C.POINTER(C.c_int)]
        L.pg_chain_run.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(PgChainStep), C.POINTER(C.c_float)]
        L.pg_chain_begin.argtypes = [vp, C.c_int]
        L.pg_chain_end.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(PgChainStep), C.POINTER(C.c_float)]
        L.pg_chain_run_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.pg_chain_steps.argtypes = [vp, C.c_int, C.c_int, C.POINTER(PgChainStep)]
        L.pg_chain_trial_xyz.argtypes = [vp, C.c_int, c_double_p, C.c_int]
        L.pg_chain_check.argtypes = [vp, C.POINTER(C.c_int)]
        L.pg_chain_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
        L.pg_chain_run_multi_io.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(C.c_uint32), c_uint8_p, C.POINTER(PgChainStep),
                                            C.POINTER(c_double_p), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.pg_measure_l2_peak.argtypes = [vp, c_double_p]
        L.pg_recompute_sk.argtypes = [vp, c_double_p]
        L.pg_recompute_sk_begin.argtypes = [vp]
        L.pg_recompute_sk_end.argtypes = [vp, c_double_p, C.POINTER(C.c_float)]
        L.pg_sk_export.argtypes = [vp, C.c_void_p]
        L.pg_sk_attach.argtypes = [vp, C.c_int, C.c_int, C.c_void_p]
        L.pg_sk_attach_local.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
        L.pg_sk_detach.argtypes = [vp]
        _LIB = L
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One system replica resident on one GPU."""

    def __init__(self, params: dict, device: int = 0, capacity_beads: int = 1024):
        self.L = lib()
        self.pb = ParamBlock(params)
        h = C.c_void_p()
        rc = self.L.pg_create(C.byref(self.pb.struct), int(device), int(capacity_beads), C.byref(h))
        if rc != 0:
            raise EngineError(f"pg_create failed ({rc}): {self.L.pg_last_error(None).decode()}")
        self.h = h
        self.mol_first = np.zeros(1, dtype=np.int32)

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.L.pg_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- set-up ------------------------------------------------------------
    def ewald_info(self) -> PgEwaldInfo:
        o = PgEwaldInfo()
        self._check(self.L.pg_get_ewald_info(self.h, C.byref(o)), "pg_get_ewald_info")
        return o

    def upload(self, xyz, q, type_ids, mol_first):
        xyz = _f64(xyz).reshape(-1, 3)
        q = _f64(q)
        t = np.ascontiguousarray(type_ids, dtype=np.int32)
        mf = np.ascontiguousarray(mol_first, dtype=np.int32)
        self.mol_first = mf.copy()
        self._check(self.L.pg_upload_system(self.h, xyz.shape[0], dptr(xyz), dptr(q), iptr(t), mf.shape[0] - 1,
                                            iptr(mf)), "pg_upload_system")

    @property
    def n(self) -> int:
        return self.L.pg_num_beads(self.h)

    def positions(self) -> np.ndarray:
        out = np.zeros((max(self.n, 1), 3), dtype=np.float64)
        self._check(self.L.pg_download_positions(self.h, dptr(out)), "pg_download_positions")
        return out[:self.n]

    def init_energy(self) -> dict:
        t = PgTotals()
        self._check(self.L.pg_init_energy(self.h, C.byref(t)), "pg_init_energy")
        return t.as_dict()

    def recompute_totals(self) -> dict:
        t = PgTotals()
        self._check(self.L.pg_recompute_totals(self.h, C.byref(t)), "pg_recompute_totals")
        return t.as_dict()

    def totals(self) -> dict:
        t = PgTotals()
        self._check(self.L.pg_get_totals(self.h, C.byref(t)), "pg_get_totals")
        return t.as_dict()

    # -- per-move ----------------------------------------------------------
    def delta_e(self, mol: int, trial_xyz, moved) -> dict:
        xyz = _f64(trial_xyz).reshape(-1, 3)
        mv = np.ascontiguousarray(moved, dtype=np.uint8)
        d = PgDelta()
        self._check(self.L.pg_delta_e(self.h, int(mol), dptr(xyz), bptr(mv), C.byref(d)), "pg_delta_e")
        return d.as_dict()

    def delta_e_raw(self, mol: int, xyz: np.ndarray, mv: np.ndarray, out: PgDelta):
        """No-allocation variant for timing loops (arrays must be contiguous f64 / u8)."""
        return self.L.pg_delta_e(self.h, mol, dptr(xyz), bptr(mv), C.byref(out))

    def delta_e_begin(self, mol: int, trial_xyz, moved) -> None:
        """Asynchronous half of delta_e (several engines driven by one thread)."""
        xyz = _f64(trial_xyz).reshape(-1, 3)
        mv = np.ascontiguousarray(moved, dtype=np.uint8)
        self._check(self.L.pg_delta_e_begin(self.h, int(mol), dptr(xyz), bptr(mv)), "pg_delta_e_begin")

    def delta_e_poll(self):
        """None while the device has not answered, else the same dict delta_e returns."""
        d = PgDelta()
        rc = self.L.pg_delta_e_poll(self.h, C.byref(d))
        if rc == 1:
            return None
        self._check(rc, "pg_delta_e_poll")
        return d.as_dict()

    def commit(self, accept: bool):
        self._check(self.L.pg_commit(self.h, int(bool(accept))), "pg_commit")

    # -- replay ------------------------------------------------------------
    def replay_upload(self, mols, offsets, us, trial_xyz, moved):
        n = len(mols)
        arr = (PgProposal * max(n, 1))()
        for i in range(n):
            arr[i].mol = int(mols[i])
            arr[i].xyz_offset = int(offsets[i])
            arr[i].u = float(us[i])
        xyz = _f64(trial_xyz).reshape(-1, 3)
        mv = np.ascontiguousarray(moved, dtype=np.uint8)
        self._check(self.L.pg_replay_upload(self.h, n, arr, xyz.shape[0], dptr(xyz), bptr(mv)), "pg_replay_upload")

    def replay_run(self, first: int, count: int):
        dE = np.zeros(max(count, 1), dtype=np.float64)
        acc = np.zeros(max(count, 1), dtype=np.uint8)
        ms = C.c_float()
        self._check(self.L.pg_replay_run(self.h, first, count, dptr(dE), bptr(acc), C.byref(ms)), "pg_replay_run")
        return dE[:count], acc[:count], ms.value

    def replay_prepare(self, first: int, count: int, with_commit: bool = True):
        self._check(self.L.pg_replay_prepare(self.h, first, count, int(with_commit)), "pg_replay_prepare")

    def replay_time_delta(self, first: int, count: int) -> float:
        ms = C.c_float()
        self._check(self.L.pg_replay_time_delta(self.h, first, count, C.byref(ms)), "pg_replay_time_delta")
        return ms.value

    # -- CBMC --------------------------------------------------------------
    # -- batched Markov-chain steps with device-side proposals --------------
    def mc_upload(self, descs, rvec=None):
        """descs: list/array of PgMoveDesc; rvec: [n][4] pivot rows."""
        n = len(descs)
        arr = (PgMoveDesc * max(n, 1))(*descs)
        rv = _f64(rvec if rvec is not None else np.zeros((0, 4))).reshape(-1, 4)
        self._mc_n = n
        self._check(self.L.pg_mc_upload(self.h, n, arr, int(rv.shape[0]), dptr(rv) if rv.size else None), "pg_mc_upload")

    def mc_run(self, first: int, count: int):
        """Returns (dE[n_done], accept[n_done], n_done, elapsed_ms)."""
        dE = np.zeros(max(count, 1))
        acc = np.zeros(max(count, 1), dtype=np.uint8)
        nd = C.c_int(0)
        ms = C.c_float(0.0)
        self._check(self.L.pg_mc_run(self.h, int(first), int(count), dptr(dE), bptr(acc), C.byref(nd), C.byref(ms)),
                    "pg_mc_run")
        return dE[:nd.value], acc[:nd.value], nd.value, float(ms.value)

    def mc_trial_xyz(self, m: int, length: int) -> np.ndarray:
        out = np.zeros((length, 3))
        self._check(self.L.pg_mc_trial_xyz(self.h, int(m), dptr(out)), "pg_mc_trial_xyz")
        return out

    # -- device-resident Markov chain (pg_chain_*) ---------------------------
    def chain_configure(self, phantom, move_size, move_prob, bond_len, vary_bond=False, gc_freq=0, cluster=1,
                        keep_trials=False, pivot_mode=0):
        c = PgChainConfig(phantom=int(phantom), gc_freq=int(gc_freq), vary_bond=int(bool(vary_bond)),
                          cluster_ctas=int(cluster), keep_trials=int(bool(keep_trials)), pivot_mode=int(pivot_mode),
                          move_size=float(move_size), bond_len=float(bond_len))
        for i in range(5):
            c.move_prob[i] = float(move_prob[i])
        self._check(self.L.pg_chain_configure(self.h, C.byref(c)), "pg_chain_configure")

    def chain_seed(self, seed: int):
        """Generator state of std::mt19937(seed) (== numpy's legacy init_genrand seeding; position 624)."""
        key = np.random.RandomState(int(seed)).get_state()[1].astype(np.uint32)
        self.chain_set_rng(key, 624)

    def chain_set_rng(self, state624, position: int):
        st = np.ascontiguousarray(state624, dtype=np.uint32)
        assert st.shape == (624,)
        self._check(self.L.pg_chain_set_rng(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32)), int(position)), "pg_chain_set_rng")

    def chain_get_rng(self):
        st = np.zeros(624, dtype=np.uint32)
        pos = C.c_int(0)
        self._check(self.L.pg_chain_get_rng(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos)), "pg_chain_get_rng")
        return st, pos.value

    def chain_run(self, max_steps: int):
        """Returns (records as a structured array with dE / mol / kind / accept / stage, stop_kind, elapsed_ms)."""
        arr = (PgChainStep * max(max_steps, 1))()
        nd, stop, ms = C.c_int(0), C.c_int(0), C.c_float(0.0)
        self._check(self.L.pg_chain_run(self.h, int(max_steps), C.byref(nd), C.byref(stop), arr, C.byref(ms)), "pg_chain_run")
        dt = np.dtype([("dE", "<f8"), ("mol", "<i4"), ("kind", "i1"), ("accept", "u1"), ("stage", "u1"), ("_pad", "u1")])
        rec = np.frombuffer(arr, dtype=dt, count=max(max_steps, 1))[:nd.value].copy()
        return rec, stop.value, float(ms.value)

    def chain_trial_xyz(self, step: int, n_beads: int) -> np.ndarray:
        out = np.zeros((n_beads, 3))
        self._check(self.L.pg_chain_trial_xyz(self.h, int(step), dptr(out), int(n_beads)), "pg_chain_trial_xyz")
        return out

    def chain_check(self) -> int:
        nb = C.c_int(0)
        self._check(self.L.pg_chain_check(self.h, C.byref(nb)), "pg_chain_check")
        if nb.value:
            raise EngineError(f"{nb.value} inconsistencies: {self.L.pg_last_error(self.h).decode()}")
        return 0

    def trial_energies(self, b1, b2, t1, q1, t2, q2, use_bead2, chain_xyz, chain_q, chain_type, current_len,
                       skip_first=-1, skip_last=-1):
        b1 = _f64(b1).reshape(-1, 3)
        nt = b1.shape[0]
        b2 = _f64(b2).reshape(-1, 3) if use_bead2 else np.zeros((nt, 3))
        n_chain = current_len * (2 if use_bead2 else 1)
        cx = _f64(chain_xyz).reshape(-1, 3) if n_chain else np.zeros((1, 3))
        cq = _f64(chain_q) if n_chain else np.zeros(1)
        ct = np.ascontiguousarray(chain_type, dtype=np.int32) if n_chain else np.zeros(1, dtype=np.int32)
        s = PgTrialSet(n_trials=nt, use_bead2=int(use_bead2), type1=int(t1), type2=int(t2), q1=float(q1), q2=float(q2),
                       current_len=int(current_len), skip_mol_first=int(skip_first), skip_mol_last=int(skip_last))
        e = np.zeros(nt)
        pe = np.zeros(nt)
        ee = np.zeros(nt)
        self._check(self.L.pg_trial_energies(self.h, C.byref(s), dptr(b1), dptr(b2), dptr(cx), dptr(cq), iptr(ct),
                                             dptr(e), dptr(pe), dptr(ee)), "pg_trial_energies")
        return e, pe, ee

    def beads_energy(self, b1, t1, q1, b2, t2, q2, use_bead2, chain_xyz, chain_q, chain_type, current_len,
                     skip_first=-1, skip_last=-1):
        """Single-trial form with the oracle's signature (tests/replay.py)."""
        e, pe, ee = self.trial_energies(np.asarray(b1).reshape(1, 3), np.asarray(b2).reshape(1, 3), t1, q1, t2, q2,
                                        use_bead2, chain_xyz, chain_q, chain_type, current_len, skip_first, skip_last)
        return float(e[0]), float(pe[0]), float(ee[0])

    def wall_force(self, phantom: int) -> np.ndarray:
        """One sample of the slab wall-force pressure sums (pg_wall_force)."""
        out = np.zeros(6)
        self._check(self.L.pg_wall_force(self.h, int(phantom), dptr(out)), "pg_wall_force")
        return out

    def vol_scaling_sample(self, phantom: int, dz: float = 1e-5) -> dict:
        """One sample of the volume-perturbation pressure estimator (pg_vol_scaling_sample): dU of a virtual
        stretch of the box along z by dz, per species pair."""
        o = PgVolSample()
        self._check(self.L.pg_vol_scaling_sample(self.h, int(phantom), float(dz), C.byref(o)), "pg_vol_scaling_sample")
        return {"el": np.array(o.el[:]), "hs": np.array(o.hs[:]), "bond": o.bond, "dipole": o.dipole, "dU": o.dU,
                "n_free": o.n_free}

    # -- GC ----------------------------------------------------------------
    def insert_molecules(self, mol_len, xyz, q, type_ids) -> dict:
        ml = np.ascontiguousarray(mol_len, dtype=np.int32)
        xyz = _f64(xyz).reshape(-1, 3)
        q = _f64(q)
        t = np.ascontiguousarray(type_ids, dtype=np.int32)
        a = PgTotals()
        self._check(self.L.pg_insert_molecules(self.h, ml.shape[0], iptr(ml), dptr(xyz), dptr(q), iptr(t),
                                               C.byref(a)), "pg_insert_molecules")
        last = int(self.mol_first[-1])
        self.mol_first = np.concatenate([self.mol_first, last + np.cumsum(ml)]).astype(np.int32)
        return a.as_dict()

    def delete_molecules(self, mf: int, ml: int) -> dict:
        r = PgTotals()
        self._check(self.L.pg_delete_molecules(self.h, int(mf), int(ml), C.byref(r)), "pg_delete_molecules")
        first = self.mol_first
        nrem = int(first[ml + 1] - first[mf])
        self.mol_first = np.concatenate([first[:mf + 1], first[ml + 2:] - nrem]).astype(np.int32)
        return r.as_dict()

    # -- reciprocal space --------------------------------------------------
    def sk_compute_slice(self, k_first: int, k_count: int, dev_ptr: int = 0):
        self._check(self.L.pg_sk_compute_slice(self.h, k_first, k_count, C.c_void_p(dev_ptr or None)),
                    "pg_sk_compute_slice")

    def sk_set(self, dev_ptr: int):
        self._check(self.L.pg_sk_set(self.h, C.c_void_p(dev_ptr)), "pg_sk_set")

    def sk_energy(self, k_first: int, k_count: int) -> float:
        e = C.c_double()
        self._check(self.L.pg_sk_energy(self.h, k_first, k_count, C.cast(C.byref(e), c_double_p)), "pg_sk_energy")
        return e.value

    def sk_download(self) -> np.ndarray:
        nk = self.ewald_info().n_k_half
        out = np.zeros((max(nk, 1), 2))
        self._check(self.L.pg_sk_download(self.h, dptr(out)), "pg_sk_download")
        return out[:nk]

    def recompute_sk(self) -> float:
        """Full S(k) recompute into the engine's own S(k) (k-sharded when peers are attached); returns the reciprocal energy."""
        e = C.c_double()
        self._check(self.L.pg_recompute_sk(self.h, C.cast(C.byref(e), c_double_p)), "pg_recompute_sk")
        return e.value

    def recompute_sk_begin(self):
        self._check(self.L.pg_recompute_sk_begin(self.h), "pg_recompute_sk_begin")

    def recompute_sk_end(self):
        e, ms = C.c_double(), C.c_float()
        self._check(self.L.pg_recompute_sk_end(self.h, C.cast(C.byref(e), c_double_p), C.byref(ms)), "pg_recompute_sk_end")
        return e.value, float(ms.value)

    def sk_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.L.pg_sk_export(self.h, buf), "pg_sk_export")
        return buf.raw

    def sk_attach(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        self._check(self.L.pg_sk_attach(self.h, int(rank), int(world), handles), "pg_sk_attach")

    def sk_attach_local(self, rank: int, peers):
        arr = (C.c_void_p * len(peers))(*[p.h for p in peers])
        self._check(self.L.pg_sk_attach_local(self.h, int(rank), len(peers), arr), "pg_sk_attach_local")

    def sk_detach(self):
        self._check(self.L.pg_sk_detach(self.h), "pg_sk_detach")

    # -- instrumentation ---------------------------------------------------
    def launch_count(self) -> int:
        return int(self.L.pg_launch_count(self.h))

    def stream(self) -> int:
        return int(self.L.pg_stream(self.h) or 0)

    def chain_counters(self, reset=False):
        """(pair configurations evaluated inside a cutoff, evaluated steps, accepted steps) since the last reset."""
        out = (C.c_uint64 * 3)()
        self._check(self.L.pg_chain_counters(self.h, out, int(bool(reset))), "pg_chain_counters")
        return int(out[0]), int(out[1]), int(out[2])

    def measure_l2_peak(self) -> float:
        g = C.c_double()
        self._check(self.L.pg_measure_l2_peak(self.h, C.cast(C.byref(g), c_double_p)), "pg_measure_l2_peak")
        return g.value

    def measure_fp64_peak(self) -> float:
        g = C.c_double()
        self._check(self.L.pg_measure_fp64_peak(self.h, C.cast(C.byref(g), c_double_p)), "pg_measure_fp64_peak")
        return g.value
