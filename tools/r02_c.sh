#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest.txt
cat gpurun_out/r02c_pytest.txt
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 1,8 --prof --check > gpurun_out/r02c_prof_S.jsonl 2>&1
cat gpurun_out/r02c_prof_S.jsonl
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 2,4,16 --replicas 74,148,296 --multi-cluster 1 > gpurun_out/r02c_probe_S.jsonl 2>&1
cat gpurun_out/r02c_probe_S.jsonl
