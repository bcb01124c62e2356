#!/usr/bin/env python3
"""Wall time of bin/plum_gpu on the four reference examples (10^5 steps, shipped sampling frequencies) with the
batched translational steps on (default) and off (PLUM_B200_BATCH=0).  Prints one JSON line per run."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import replay  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for name in ("bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"):
    for batch in ("2", "1", "0"):
        t0 = time.perf_counter()
        lines = replay.run_plum_ref(replay.golden_example_dir(name), steps, 1, xyz=False, binary=replay.PLUM_GPU,
                                    extra_env={"PLUM_B200_BATCH": batch})
        dt = time.perf_counter() - t0
        print(json.dumps({"example": name, "PLUM_B200_BATCH": batch, "steps": steps, "wall_s": round(dt, 2),
                          "us_per_step": round(dt / steps * 1e6, 2)}), flush=True)
