#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_chain_gpu.py -x -q -m gpu -k "oracle or totals or init or parity or random_moves" -s 2>&1 | grep -v "^$" | tail -12
timeout -k 10 1200 ncu --set full --clock-control none -k regex:k_ -f -o /tmp/prof_r02n_all python tools/profile_kernels.py > gpurun_out/r02n_ncu_all.log 2>&1; tail -3 gpurun_out/r02n_ncu_all.log
ncu -i /tmp/prof_r02n_all.ncu-rep --page raw --csv > gpurun_out/prof_r02n_all_raw.csv 2>/dev/null
python tools/kernels_summary.py gpurun_out/prof_r02n_all_raw.csv gpurun_out/r02n_kernels.json | tail -50
timeout 600 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; tail -c 1500 gpurun_out/r02n_bench.json; tail -5 gpurun_out/r02n_bench.err
