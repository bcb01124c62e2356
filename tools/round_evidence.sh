#!/bin/bash
# One gpurun call that regenerates the round's measured evidence (tests, bench lines, ncu captures).
# usage: tools/round_evidence.sh r01   -> gpurun_out/{pytest_gpu,bench_n1,bench_ref,launches,prof_*}_<tag>.*
tag=${1:-r01}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log; fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; tail -c 600 gpurun_out/bench_n1_$tag.json
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 400 gpurun_out/bench_ref_$tag.json
# throughput-mode grid shapes (what the headline runs: whole-tile pair CTAs, 3 CTA slots per SM), one replica so that ncu can serialise
PLUM_B200_TILE_CTAS=1 PLUM_B200_CTAS_PER_SM=3 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --moves-per-step 64 --no-cpu-baseline --replicas-per-gpu 1 > gpurun_out/ncu_bench1_$tag.log 2>&1
PLUM_B200_TILE_CTAS=1 PLUM_B200_CTAS_PER_SM=3 ncu --set full --clock-control none --import-source on -k regex:k_move -s 400 -c 8 -f -o gpurun_out/prof_${tag}_kmove \
    python bench.py --steps 1 --warmup 3 --moves-per-step 64 --no-cpu-baseline --replicas-per-gpu 1 > gpurun_out/ncu_bench2_$tag.log 2>&1
# the single-chain split of the same kernel (finest one-wave grid)
ncu --set full --clock-control none -k regex:k_move -s 400 -c 4 -f -o /tmp/prof_${tag}_kmove_single \
    python bench.py --steps 1 --warmup 3 --moves-per-step 64 --no-cpu-baseline --replicas-per-gpu 1 > gpurun_out/ncu_bench3_$tag.log 2>&1
ncu -i /tmp/prof_${tag}_kmove_single.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_kmove_single_raw.csv 2>/dev/null
# every kernel of the path; only the raw-metric CSV travels back (gpurun_out is capped at 64 MiB)
ncu --set full --clock-control none -k regex:k_ -f -o /tmp/prof_${tag}_all \
    python tools/profile_kernels.py > gpurun_out/ncu_all_$tag.log 2>&1; tail -2 gpurun_out/ncu_all_$tag.log
ncu -i /tmp/prof_${tag}_all.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_all_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
