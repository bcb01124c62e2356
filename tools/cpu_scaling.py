#!/usr/bin/env python3
"""SURVEY.md §8(d) CPU baseline (2a): the REAL reference binary (oracle/_ref/plum_ref, one core) on down-scaled cuts
of the synthetic system S — same box L = 200 and alpha (hence the same cutoffs and K = 3574), fewer chains — and a fit
t_move = a * n_moved * N + b per move kind, extrapolated to N = 22 000 (which plum_ref cannot hold: its std::map
caches need 70-115 GB).  Runs where oracle/_ref/plum_ref exists (the build container); prints one JSON line per size
and a final line with the fit.  usage: tools/cpu_scaling.py [chains ...]   (default 12 25 50)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import replay  # noqa: E402
from plum_b200 import synth  # noqa: E402

chains = [int(x) for x in sys.argv[1:]] or [12, 25, 50]
STEPS = 60
rows = []
for nc in chains:
    sysm = synth.make_system(n_chains=nc, chain_len=100, charged_every=10)
    with tempfile.TemporaryDirectory(prefix="cpu_scaling_") as d:
        synth.write_inputs(d, sysm, n_steps=STEPS, alpha=0.004, spring=False)
        t0 = time.perf_counter()
        replay.run_plum_ref(d, 0, 1, xyz=False)                 # initialisation only
        t_init = time.perf_counter() - t0
        t0 = time.perf_counter()
        lines = replay.run_plum_ref(d, STEPS, 1, xyz=False)
        t_run = time.perf_counter() - t0 - t_init
    T = [ln.split() for ln in lines if ln.startswith("T ")]
    n_ion = sum(1 for t in T if t[2] == "0")
    n_chain = len(T) - n_ion
    row = {"chains": nc, "N": sysm.n, "init_s": round(t_init, 2), "moves_s": round(t_run, 2), "moves": len(T),
           "ion_moves": n_ion, "chain_moves": n_chain, "s_per_move": round(t_run / max(len(T), 1), 4)}
    rows.append(row)
    print(json.dumps(row), flush=True)
# per move the reference does n_moved * N pair evaluations of (125 images + 3574 cos): time ~ (n_ion + 100 n_chain) * N
x = np.array([(r["ion_moves"] + 100.0 * r["chain_moves"]) * r["N"] for r in rows])
y = np.array([r["moves_s"] for r in rows])
a = float((x @ y) / (x @ x))
N = 22000
p_ion = 0.5
t_move = a * N * (p_ion * 1 + (1 - p_ion) * 100)
print(json.dumps({"fit_s_per_moved_bead_partner": a, "extrapolated_s_per_move_at_N22000": t_move,
                  "extrapolated_moves_per_s_at_N22000": 1.0 / t_move,
                  "note": "EXTRAPOLATED from the sizes above (plum_ref cannot hold N = 22000); one core; move mix 0.5 ion / 0.5 chain"}))
