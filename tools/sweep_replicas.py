#!/usr/bin/env python3
"""bench.py at several replicas-per-GPU settings: value / e2e / roofline per setting."""
import json, subprocess, sys
for R in [int(x) for x in (sys.argv[1:] or ["4", "8", "12", "16"])]:
    o = subprocess.run([sys.executable, "bench.py", "--replicas-per-gpu", str(R), "--no-cpu-baseline", "--no-single"],
                       capture_output=True, text=True)
    try:
        b = json.loads(o.stdout.strip().splitlines()[-1])
        print(f"R={R}: value {b['value']:.0f} e2e {b['e2e']['value']:.0f} frac {b['roofline']['frac']:.3f} "
              f"avg_launch_us {b['roofline']['avg_launch_us']:.2f} ms/step {b['ms_per_step']:.2f}", flush=True)
    except Exception as e:   # noqa: BLE001
        print(f"R={R}: failed {e} {o.stderr[-400:]}", flush=True)
