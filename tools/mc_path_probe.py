#!/usr/bin/env python3
"""One Markov chain, two ways through the C ABI, wall-clock per step on the host (plum_b200/host/mc_bench.cc):
  per-move   pg_delta_e_begin/_poll + pg_commit, trial coordinates built on the host
  batched    pg_mc_upload/_begin/_end, device-side proposals + Metropolis test + commit
for the reference examples' systems and the 22 000-bead synthetic system S.  Prints one JSON line per system
(also the fraction of steps that end in dE >= 1e8, which is what limits the batch length)."""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import replay  # noqa: E402
from plum_b200 import mcgen, synth  # noqa: E402
from plum_b200.engine import Engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000


def system(name):
    if name == "S":
        r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    else:
        r, s, types, params = replay.load_golden(name)
    return r, s, types, params


for name in ("bulk_nvt", "confined_nvt", "synth_spring", "S"):
    r, s, types, params = system(name)
    out = {"system": name, "beads": s.n, "steps": N}
    for mode in ("per_move", "batched"):
        eng = Engine(params, device=0, capacity_beads=s.n)
        eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
        eng.init_energy()
        chain = mcgen.NativeChain(eng, r, s, 3, record_moves=N)
        warm = min(2000, N)
        (chain.run_per_move if mode == "per_move" else chain.run_batched)(warm, *(() if mode == "per_move" else (1024,)))
        l0 = eng.launch_count()
        wall = chain.run_per_move(N) if mode == "per_move" else chain.run_batched(N, 1024)
        out[mode + "_us_per_step"] = round(wall / N * 1e6, 2)
        out[mode + "_launches_per_step"] = round((eng.launch_count() - l0) / N, 2)
        if mode == "per_move":
            out["overlap_fraction"] = round(float(np.mean(chain.rec_dE >= 1e8)), 4)
            out["accept_ratio"] = round(float(chain.rec_acc.mean()), 4)
            ref = (chain.rec_mol.copy(), chain.rec_acc.copy(), chain.rec_dE.copy())
        else:
            out["same_chain"] = bool(np.array_equal(ref[0], chain.rec_mol) and np.array_equal(ref[1], chain.rec_acc) and
                                     np.array_equal(ref[2], chain.rec_dE))
        chain.close()
        eng.close()
    print(json.dumps(out), flush=True)
