#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_chain_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-recompute --no-examples > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; tail -5 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02q_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','resident_matches_e2e')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['replicas_per_gpu'])
print(json.dumps(d['single_chain'])[:300])
PY
