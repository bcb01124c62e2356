#!/bin/bash
# Throughput-mode CTA granularity of k_move chain moves: tiles per pair CTA x helper CTAs (n_sm / div).
# usage (on the GPU box): tools/sweep_tile_ctas.sh  -> one line per setting: tiles_per_cta helpers_div value e2e
for t in 1 2 4; do for d in 1 2 4; do
  PLUM_B200_TILE_CTAS=1 PLUM_B200_TILES_PER_CTA=$t PLUM_B200_HELPERS_DIV=$d timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single > /tmp/sw.json 2>/dev/null
  python -c "
import json;d=json.load(open('/tmp/sw.json'));print($t, $d, round(d['value']), round(d['e2e_per_move']['value']), round(d['roofline']['avg_launch_us'],1), d['replay_matches_e2e'])"
done; done
