#!/bin/bash
# Short gpurun call for the end of a round (few GPU-minutes left): the newest tests first, then a short bench line,
# then as much of the remaining GPU suite as the time allows.  usage: tools/final_check.sh <tag>
tag=${1:-r01e}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_trajectory_gpu.py -x -q -m gpu -k "1-2 or 0-3" > gpurun_out/pytest_seeds_$tag.log 2>&1; tail -3 gpurun_out/pytest_seeds_$tag.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline > gpurun_out/bench_short_$tag.json 2> gpurun_out/bench_short_$tag.err; tail -c 1500 gpurun_out/bench_short_$tag.json
timeout 300 python -m pytest tests -x -q -m gpu --deselect tests/test_trajectory_gpu.py > gpurun_out/pytest_rest_$tag.log 2>&1; tail -3 gpurun_out/pytest_rest_$tag.log
# what the e2e leg gives with the few host threads per GPU an 8-GPU box leaves (host cores / 8 - 2)
timeout 120 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute --host-threads 6 > gpurun_out/bench_ht6_$tag.json 2> gpurun_out/bench_ht6_$tag.err; tail -c 300 gpurun_out/bench_ht6_$tag.json
