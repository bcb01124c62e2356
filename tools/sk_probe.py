#!/usr/bin/env python3
"""Full S(k) recompute of S / S-full a few times (for ncu launch lists and timing)."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from plum_b200 import synth, sharded
from plum_b200.engine import Engine
for name, loader in (("S", synth.load), ("S_full", synth.load_full)):
    r, s, types, params = loader(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    e = Engine(params, device=0, capacity_beads=s.n)
    e.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    t = e.init_energy()
    print(name, json.dumps(sharded.time_fused_recompute(e, 0, 1, t["recip"], reps=int(os.environ.get("REPS", "30")))))
    e.close()
