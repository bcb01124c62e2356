#!/bin/bash
# A/B on one GPU box (scratch copy of the repo): baseline first, then every variants/<name>/libplum_b200.so swapped in.
# Per build: a short bench line and the GPU tests that run k_move<true> (parity vs the oracle and the reference traces).
#   bash tools/ab_run.sh <tag> <name> [<name> ...]      -> gpurun_out/ab_<tag>_<name>.{json,log}
tag=$1; shift
mkdir -p gpurun_out
cp plum_b200/libplum_b200.so /tmp/base_libplum_b200.so
run_one() {
  name=$1
  timeout 60 python bench.py --steps 3 --warmup 3 --no-single --no-cpu-baseline --no-recompute \
      > "gpurun_out/ab_${tag}_${name}.json" 2> "gpurun_out/ab_${tag}_${name}.err"
  python - "$name" "gpurun_out/ab_${tag}_${name}.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:40s} value {d['value']:10.0f}  e2e {d['e2e']['value']:10.0f}  ms/step {d['ms_per_step']:.2f}  "
          f"replay_matches_e2e {d['replay_matches_e2e']}")
except Exception as e:
    print(sys.argv[1], "bench failed:", e)
PY
  timeout 120 python -m pytest tests -x -q -m gpu --deselect tests/test_trajectory_gpu.py \
      -k "synth or full_size or s_full or spring or per_move" > "gpurun_out/ab_${tag}_${name}.log" 2>&1
  echo "    tests: $(tail -1 gpurun_out/ab_${tag}_${name}.log)"
}
run_one baseline
for v in "$@"; do
  cp "variants/$v/libplum_b200.so" plum_b200/libplum_b200.so
  run_one "$(echo "$v" | tr '+' '_')"
done
cp /tmp/base_libplum_b200.so plum_b200/libplum_b200.so
