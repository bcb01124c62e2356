#!/bin/bash
# Round-2 evidence, ONE gpurun call on one GPU:  bash tools/round2_evidence.sh r02
# One `ncu --set full` capture of k_chain (single chain, one CTA, 300 steps of S) and its summary (copied to
# profiles/<tag>_ncu_summary.json on the box so that the bench run below reads the counters of THIS build), the GPU suite,
# smoke, bench (both arms) and the ncu launch list of the bench command -> gpurun_out/<tag>_* ; copy what is to be judged
# into profiles/.
tag=${1:-r02}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_${tag}_kchain \
    python tools/chain_probe.py --system S --steps 300 --clusters 1 > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_chain_summary.py gpurun_out/prof_${tag}_kchain.ncu-rep 300 gpurun_out/${tag}_ncu_summary.json | head -c 400
# the whole fleet (two chains per SM, the 448-thread build) is ONE launch: its capture gives device-wide issue / FP64-pipe / L2 percentages
timeout -k 10 900 ncu --set full --clock-control none -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_${tag}_kchain_fleet \
    python tools/chain_probe.py --system S --steps 100 --clusters "" --replicas 296 > gpurun_out/${tag}_ncu_fleet.log 2>&1
python tools/ncu_chain_summary.py gpurun_out/prof_${tag}_kchain_fleet.ncu-rep 29600 gpurun_out/${tag}_ncu_summary.json k_chain_fleet | tail -c 600
cp gpurun_out/${tag}_ncu_summary.json profiles/${tag}_ncu_summary.json
timeout -k 10 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/${tag}_pytest_gpu.txt
cat gpurun_out/${tag}_pytest_gpu.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -2 gpurun_out/${tag}_smoke.txt
timeout -k 10 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 600 gpurun_out/${tag}_bench_n1.json
timeout -k 10 900 python bench.py --impl reference > gpurun_out/${tag}_bench_ref_n1.json 2> gpurun_out/${tag}_bench_ref_n1.err; tail -c 400 gpurun_out/${tag}_bench_ref_n1.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --moves-per-step 128 --no-cpu-baseline --no-single --no-recompute --no-examples --replicas-per-gpu 16 > gpurun_out/${tag}_launches_bench.log 2>&1
# every kernel of the path once or twice (tools/profile_kernels.py) -> per-kernel rows
timeout -k 10 1200 ncu --set full --clock-control none -k regex:k_ -f -o /tmp/prof_${tag}_all python tools/profile_kernels.py > gpurun_out/${tag}_ncu_all.log 2>&1
ncu -i /tmp/prof_${tag}_all.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_all_raw.csv 2>/dev/null
python tools/kernels_summary.py gpurun_out/prof_${tag}_all_raw.csv gpurun_out/${tag}_kernels.json | tail -40
