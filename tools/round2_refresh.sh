#!/bin/bash
# After a change of k_chain only: the two ncu captures, the chain tests, the bench line.  bash tools/round2_refresh.sh r02
tag=${1:-r02}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_${tag}_kchain \
    python tools/chain_probe.py --system S --steps 300 --clusters 1 > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_chain_summary.py gpurun_out/prof_${tag}_kchain.ncu-rep 300 gpurun_out/${tag}_ncu_summary.json | head -c 300
timeout -k 10 900 ncu --set full --clock-control none -k regex:k_chain -s 1 -c 1 -f -o gpurun_out/prof_${tag}_kchain_fleet \
    python tools/chain_probe.py --system S --steps 100 --clusters "" --replicas 296 > gpurun_out/${tag}_ncu_fleet.log 2>&1
python tools/ncu_chain_summary.py gpurun_out/prof_${tag}_kchain_fleet.ncu-rep 29600 gpurun_out/${tag}_ncu_summary.json k_chain_fleet | tail -c 500
cp gpurun_out/${tag}_ncu_summary.json profiles/${tag}_ncu_summary.json
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -2 gpurun_out/${tag}_smoke.txt
timeout -k 10 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 300 gpurun_out/${tag}_bench_n1.json; tail -3 gpurun_out/${tag}_bench_n1.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --moves-per-step 128 --no-cpu-baseline --no-single --no-recompute --no-examples --replicas-per-gpu 16 > gpurun_out/${tag}_launches_bench.log 2>&1
