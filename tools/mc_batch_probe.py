#!/usr/bin/env python3
"""Where a batch's time goes: for batch lengths B, wall time of pg_mc_upload, wall time of pg_mc_run and the
CUDA-event time of the batch on the device, per done step (medians over repeats), on one system."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import replay  # noqa: E402
from plum_b200 import mcgen, synth  # noqa: E402
from plum_b200.engine import Engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bulk_nvt"
if name == "S":
    r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
else:
    r, s, types, params = replay.load_golden(name)
eng = Engine(params, device=0, capacity_beads=s.n)
eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
eng.init_energy()
g = mcgen.Generator.for_run(r, s.mol_first, 9)
for B in (8, 32, 128, 512):
    rows = []
    for rep in range(12):
        descs, rv, n_rows = [], [], 0
        while len(descs) < B:
            kind, d, rvv = g.next()
            if kind < 0:
                continue
            d.rv_offset = n_rows
            descs.append(d)
            rv.append(rvv)
            n_rows += rvv.shape[0]
        t0 = time.perf_counter()
        eng.mc_upload(descs, np.concatenate(rv) if n_rows else None)
        t1 = time.perf_counter()
        l0 = eng.launch_count()
        dE, acc, n_done, ms = eng.mc_run(0, B)
        t2 = time.perf_counter()
        rows.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6, ms * 1e3, n_done, eng.launch_count() - l0))
    a = np.array(rows[2:])
    full = a[a[:, 3] == B]
    print(json.dumps({"system": name, "B": B, "upload_us": round(float(np.median(a[:, 0])), 1),
                      "run_wall_us": round(float(np.median(a[:, 1])), 1), "device_us": round(float(np.median(a[:, 2])), 1),
                      "median_done": float(np.median(a[:, 3])), "launches": float(np.median(a[:, 4])),
                      "device_us_per_step_full_batches": round(float(np.median(full[:, 2] / B)), 2) if len(full) else None,
                      "wall_us_per_step_full_batches": round(float(np.median((full[:, 0] + full[:, 1]) / B)), 2) if len(full) else None}),
          flush=True)
