#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -6
timeout -k 10 600 python -m pytest tests/test_chain_gpu.py tests/test_mc_gpu.py -x -q 2>&1 | tail -4
timeout -k 10 600 python -m pytest tests/test_trajectory_gpu.py -x -q -k "cut_of_S or (confined_nvt and 2-1)" 2>&1 | tail -4
timeout -k 10 300 python - <<'PY'
import os, sys, json, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from plum_b200 import synth, sharded
from plum_b200.engine import Engine
for name, loader in (("S", synth.load), ("S_full", synth.load_full)):
    r,s,types,params = loader(cache_dir="gpurun_out/cache")
    e = Engine(params, device=0, capacity_beads=s.n)
    e.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    t0=time.perf_counter(); t = e.init_energy(); ti=time.perf_counter()-t0
    print(name, "init_energy s", ti, json.dumps(sharded.time_fused_recompute(e, 0, 1, t["recip"])))
    e.close()
PY
