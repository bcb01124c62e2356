#!/bin/bash
# where the wall time of bin/plum_gpu goes on the reference examples (PLUM_B200_PROFILE=1 facade timers)
for ex in ${@:-bulk_nvt confined_nvt bulk_muvt confined_muvt}; do
  d=$(mktemp -d); cp tests/golden/examples/$ex/* $d/; sed -i 's/^s1_total_simulation_steps .*/s1_total_simulation_steps 100000/' $d/run.in
  echo "== $ex"
  ( cd $d; PLUM_SEED=1 PLUM_B200_PROFILE=1 bash -c "time $GRAFT_REPO_ROOT/bin/plum_gpu < run.in > run.log" 2>&1 | grep -E "profile|real" )
  rm -rf $d
done
