#!/bin/bash
# bash tools/bench_n2.sh <n_gpus> <tag>: the driver's multi-GPU launch of bench.py, short
n=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --no-examples > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "rc $?"; tail -c 800 gpurun_out/${tag}_bench_n${n}.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${tag}_bench_n${n}.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
print(json.dumps(d['sharded_recompute'])[:900])
PY
