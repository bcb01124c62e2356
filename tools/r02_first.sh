#!/bin/bash
# round 2, first GPU call: parity of the chain kernel + first timings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout -k 10 600 python -m pytest tests/test_chain_gpu.py -x -q -k "spring or interleaves" 2>&1 | tail -30 > gpurun_out/r02a_pytest1.txt
cat gpurun_out/r02a_pytest1.txt
timeout -k 10 300 python tools/chain_probe.py --system synth_cut --steps 500 --clusters 1,2,8 --check > gpurun_out/r02a_probe_cut.jsonl 2>&1
cat gpurun_out/r02a_probe_cut.jsonl
timeout -k 10 600 python -m pytest tests/test_chain_gpu.py -x -q -k "full_size or many_replicas" 2>&1 | tail -30 > gpurun_out/r02a_pytest2.txt
cat gpurun_out/r02a_pytest2.txt
timeout -k 10 600 python tools/chain_probe.py --system S --steps 2000 --clusters 1,2,4,8,16 --check --replicas 16,74,148 --multi-cluster 1,2 > gpurun_out/r02a_probe_S.jsonl 2>&1
cat gpurun_out/r02a_probe_S.jsonl
timeout -k 10 300 compute-sanitizer --tool memcheck python tools/chain_probe.py --system synth_spring --steps 60 --clusters 1,2 2>&1 | tail -25 > gpurun_out/r02a_memcheck.txt
tail -8 gpurun_out/r02a_memcheck.txt
