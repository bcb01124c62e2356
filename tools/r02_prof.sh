#!/bin/bash
# round 2: where does k_chain's time go?  phase stamps, skip experiments, one ncu --set full capture
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 300 python tools/chain_probe.py --system S --steps 1000 --clusters 1,8 --prof > gpurun_out/r02b_prof_S.jsonl 2>&1
cat gpurun_out/r02b_prof_S.jsonl
timeout -k 10 300 python tools/chain_probe.py --system S --steps 1000 --clusters 1 --skip 1,2,4,8,32,63 > gpurun_out/r02b_skip_S.jsonl 2>&1
cat gpurun_out/r02b_skip_S.jsonl
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 1 -c 1 -o gpurun_out/prof_r02b_kchain python tools/chain_probe.py --system S --steps 300 --clusters 1 > gpurun_out/r02b_ncu.log 2>&1
tail -3 gpurun_out/r02b_ncu.log
ls -la gpurun_out/*.ncu-rep
