#!/bin/bash
# ~3 GPU-minutes: GPU suite without the 10^5-step trajectories on the new build, a short bench line, and the same bench
# with the 80-register / 6-CTAs-per-SM build of k_move<true> swapped in (variants/, built with -DMV_FAST_OCC=6).
tag=${1:-r01f}
mkdir -p gpurun_out
timeout 110 python -m pytest tests -x -q -m gpu --deselect tests/test_trajectory_gpu.py > gpurun_out/pytest_rest_$tag.log 2>&1; tail -2 gpurun_out/pytest_rest_$tag.log
timeout 60 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute > gpurun_out/bench_short_$tag.json 2> gpurun_out/bench_short_$tag.err; tail -c 200 gpurun_out/bench_short_$tag.json
# (on the box's scratch copy only) nvcc ... -DMV_FAST_OCC=6 -o variants/libplum_b200_occ6.so plum_b200/csrc/pg_engine.cu beforehand
cp variants/libplum_b200_occ6.so plum_b200/libplum_b200.so
timeout 60 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute > gpurun_out/bench_occ6_$tag.json 2> gpurun_out/bench_occ6_$tag.err; tail -c 200 gpurun_out/bench_occ6_$tag.json
