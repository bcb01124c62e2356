#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-examples > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err
tail -c 1500 gpurun_out/r02l_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print(json.dumps(d['sharded_recompute']))
print(json.dumps(d['single_chain'])[:400])
PY
