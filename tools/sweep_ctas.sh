#!/bin/bash
# sweep CTAs-per-SM used by one k_move grid, single replica and 8 replicas per GPU
for c in 4 3 2; do for R in 1 8; do
  PLUM_B200_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-single --replicas-per-gpu $R > /tmp/sweep.json
  python - "$c" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep.json"))
print("ctas/sm", sys.argv[1], "R", d["config"]["replicas_per_gpu"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      d["replay_matches_e2e"], "frac", round(d["roofline"]["frac"], 3), "avg_launch_us", round(d["roofline"]["avg_launch_us"], 2))
PY
done; done
