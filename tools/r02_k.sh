#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 200 python -m pytest tests/test_engine_gpu.py -x -q -k "fused or drift" 2>&1 | tail -15
timeout -k 10 200 python tools/sk_probe.py
