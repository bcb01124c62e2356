#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 600 python -m pytest tests/test_trajectory_gpu.py -x -q -k "cut_of_S" 2>&1 | tail -12
timeout -k 10 600 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -4
timeout -k 10 900 python bench.py --no-recompute --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 2000 gpurun_out/r02h_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(json.dumps(d['single_chain'])[:900])
print(json.dumps(d['examples']))
PY
