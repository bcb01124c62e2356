#!/bin/bash
# A/B of pg_chain_config.pivot_mode on S: one chain on a cluster of 16 CTAs, and the fleet of two chains per SM.
#   gpurun -- 'bash tools/pivot_mode_ab.sh'   -> gpurun_out/pivot_mode_ab.jsonl
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/pivot_mode_ab.jsonl
for pm in 0 2 1; do
  timeout 600 python tools/chain_probe.py --system S --steps 2000 --clusters 16 --pivot-mode $pm --prof 2>&1 | grep '^{' | sed "s/^{/{\"pivot_mode\": $pm, /" | tee -a gpurun_out/pivot_mode_ab.jsonl | cut -c1-400
  timeout 600 python tools/chain_probe.py --system S --steps 1000 --clusters "" --replicas 148,296 --pivot-mode $pm 2>&1 | grep '^{' | sed "s/^{/{\"pivot_mode\": $pm, /" | tee -a gpurun_out/pivot_mode_ab.jsonl
done
