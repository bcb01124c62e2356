#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02i_pytest.txt; cat gpurun_out/r02i_pytest.txt
timeout -k 10 900 python bench.py --no-recompute --no-cpu-baseline --no-single > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 1500 gpurun_out/r02i_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(json.dumps(d['examples']))
PY
