#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_trajectory_gpu.py tests/test_chain_gpu.py -x -q -m gpu -k "vol or crank or cut_of_S or multi or replicas" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-recompute > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; tail -5 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['examples'].items():
    print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a not in ('facade_sites','timing_note')})
print(json.dumps(d['single_chain'])[:600])
PY
