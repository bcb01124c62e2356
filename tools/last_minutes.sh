#!/bin/bash
# ~100 GPU-seconds for a last A/B at the end of a round: a short bench line on the current build, the GPU tests that run
# k_move<true>, then as much of the rest of the suite (without the 10^5-step trajectories) as the time allows.
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 40 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute > gpurun_out/bench_short_$tag.json 2> gpurun_out/bench_short_$tag.err; tail -c 250 gpurun_out/bench_short_$tag.json
timeout 50 python -m pytest tests -x -q -m gpu --deselect tests/test_trajectory_gpu.py -k "synth or full_size or s_full or spring or per_move" > gpurun_out/pytest_fast_$tag.log 2>&1; tail -2 gpurun_out/pytest_fast_$tag.log
timeout 80 python -m pytest tests -x -q -m gpu --deselect tests/test_trajectory_gpu.py -k "not (synth or full_size or s_full or spring or per_move)" > gpurun_out/pytest_rest_$tag.log 2>&1; tail -2 gpurun_out/pytest_rest_$tag.log
