#!/usr/bin/env python3
"""Warm per-kernel timing breakdown on the synthetic S system (run under gpurun)."""
import os, sys, time, ctypes as C
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from plum_b200 import synth
from plum_b200.engine import Engine
from plum_b200._abi import PgDelta

r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
eng = Engine(params, device=0, capacity_beads=s.n)
ids = types.ids(s.symbol)
eng.upload(s.xyz, s.q, ids, s.mol_first)
t0 = time.perf_counter(); eng.init_energy(); print("init_energy s", time.perf_counter() - t0)
rng = np.random.default_rng(0)
chains = [m for m in range(s.n_mol) if s.mol_first[m+1]-s.mol_first[m] > 1]
ions = [m for m in range(s.n_mol) if s.mol_first[m+1]-s.mol_first[m] == 1]
def proposals(mols):
    offs, xyz, mv, off = [], [], [], 0
    for m in mols:
        f, l = s.mol_first[m], s.mol_first[m+1]
        xyz.append(s.xyz[f:l] + rng.normal(scale=0.3, size=(l-f, 3)))
        mv.append(np.ones(l-f, dtype=np.uint8)); offs.append(off); off += l-f
    return offs, np.concatenate(xyz), np.concatenate(mv)
for name, pool in (("ion", ions), ("chain", chains)):
    mols = [int(rng.choice(pool)) for _ in range(512)]
    offs, xyz, mv = proposals(mols)
    eng.replay_upload(mols, offs, [2.0]*len(mols), xyz, mv)   # u=2: never accepted
    eng.replay_time_delta(0, 128)
    ms = eng.replay_time_delta(0, 512)
    print(f"{name}: k_delta back-to-back avg {ms*1e3/512:.2f} us")
    dE, acc, ms2 = eng.replay_run(0, 512)
    dE, acc, ms2 = eng.replay_run(0, 512)
    print(f"{name}: k_delta+k_commit back-to-back avg {ms2*1e3/512:.2f} us")
    # host-synchronous path
    d = PgDelta()
    for rep in range(2):
        t0 = time.perf_counter()
        for i, m in enumerate(mols[:256]):
            f, l = s.mol_first[m], s.mol_first[m+1]
            x = np.ascontiguousarray(xyz[offs[i]:offs[i]+(l-f)]); v = np.ascontiguousarray(mv[offs[i]:offs[i]+(l-f)])
            eng.delta_e_raw(m, x, v, d); eng.L.pg_commit(eng.h, 0)
        dt = time.perf_counter() - t0
    print(f"{name}: pg_delta_e+pg_commit (python ctypes caller) avg {dt*1e6/256:.2f} us")
print("fp64 peak GF", eng.measure_fp64_peak())
