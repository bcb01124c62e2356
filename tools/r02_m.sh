#!/bin/bash
# A/B: 512-thread k_chain (1 chain per SM) vs the 256-thread build (2 chains per SM), whole fleet in one launch.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cp plum_b200/libplum_b200.so /tmp/base.so
echo "== base (512 threads)"; timeout 600 python tools/chain_probe.py --system S --steps 1000 --clusters "" --replicas 148,296 2>&1 | tee gpurun_out/r02m_base.jsonl
cp variants/t256/libplum_b200.so plum_b200/libplum_b200.so
echo "== t256 (256 threads, 2 CTAs/SM)"; timeout 600 python tools/chain_probe.py --system S --steps 1000 --clusters "" --replicas 148,296,592 2>&1 | tee gpurun_out/r02m_t256.jsonl
timeout 300 python -m pytest tests/test_chain_gpu.py -x -q -m gpu -k "1320 or spring or crank" 2>&1 | tail -3
cp /tmp/base.so plum_b200/libplum_b200.so
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_mc_gpu.py tests/test_trajectory_gpu.py -x -q -m gpu -k "crank" 2>&1 | tail -5
