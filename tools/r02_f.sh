#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r02f_pytest.txt
cat gpurun_out/r02f_pytest.txt
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 1,8,16 --prof --check > gpurun_out/r02f_prof_S.jsonl 2>&1
cat gpurun_out/r02f_prof_S.jsonl
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 4,8,16 --pivot-mode 1 --prof > gpurun_out/r02f_prof_S_pm1.jsonl 2>&1
cat gpurun_out/r02f_prof_S_pm1.jsonl
timeout -k 10 300 python tools/chain_probe.py --system S --steps 2000 --clusters 16 --replicas 148,296 --multi-cluster 1 > gpurun_out/r02f_probe_S.jsonl 2>&1
tail -2 gpurun_out/r02f_probe_S.jsonl
