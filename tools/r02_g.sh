#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 1,16 --replicas 148,296 --multi-cluster 1 > gpurun_out/r02g_probe_S.jsonl 2>&1
cat gpurun_out/r02g_probe_S.jsonl
timeout -k 10 900 python bench.py --no-recompute > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -c 3000 gpurun_out/r02g_bench.err; cat gpurun_out/r02g_bench.json
