#!/usr/bin/env python3
"""Per-call latency of pg_delta_e + pg_commit on the reference examples (generic k_move path)."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import replay
from plum_b200.engine import Engine
from plum_b200._abi import PgDelta
for name in ("bulk_nvt", "confined_nvt"):
    r, s, types, params = replay.load_golden(name)
    eng = Engine(params, device=0, capacity_beads=s.n + 64)
    eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first); eng.init_energy()
    rng = np.random.default_rng(0)
    chains = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m+1]-s.mol_first[m] > 1]
    ions = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m+1]-s.mol_first[m] == 1]
    d = PgDelta()
    for label, pool in (("ion", ions), ("chain", chains)):
        moves = []
        for _ in range(300):
            m = int(rng.choice(pool)); f, l = s.mol_first[m], s.mol_first[m+1]
            moves.append((m, np.ascontiguousarray(s.xyz[f:l] + rng.normal(scale=0.3, size=(l-f, 3))), np.ones(l-f, dtype=np.uint8)))
        for rep in range(2):
            t0 = time.perf_counter()
            for m, x, v in moves:
                eng.delta_e_raw(m, x, v, d); eng.L.pg_commit(eng.h, 0)
            dt = time.perf_counter() - t0
        print(f"{name} {label}: {dt*1e6/len(moves):.1f} us per pg_delta_e+pg_commit (python caller), N={s.n}", flush=True)
    eng.close()
