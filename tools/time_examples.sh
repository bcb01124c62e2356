#!/bin/bash
# wall time of Plum's driver on the four reference examples (10^5 steps, shipped sampling), B200 facade
for ex in bulk_nvt confined_nvt bulk_muvt confined_muvt; do
  d=$(mktemp -d); cp tests/golden/examples/$ex/* $d/; sed -i 's/^s1_total_simulation_steps .*/s1_total_simulation_steps 100000/' $d/run.in
  ( cd $d; s=$(date +%s.%N); PLUM_SEED=1 $GRAFT_REPO_ROOT/bin/plum_gpu < run.in > run.log; e=$(date +%s.%N); echo "$ex plum_gpu 1e5 steps: $(echo "$e - $s" | bc) s" )
  sed -i 's/^s1_total_simulation_steps .*/s1_total_simulation_steps 0/' $d/run.in
  ( cd $d; s=$(date +%s.%N); PLUM_SEED=1 $GRAFT_REPO_ROOT/bin/plum_gpu < run.in > run.log; e=$(date +%s.%N); echo "$ex plum_gpu init only: $(echo "$e - $s" | bc) s" )
  if [ -x oracle/_ref/plum_ref ]; then
    sed -i 's/^s1_total_simulation_steps .*/s1_total_simulation_steps 2000/' $d/run.in
    ( cd $d; s=$(date +%s.%N); PLUM_SEED=1 $GRAFT_REPO_ROOT/oracle/_ref/plum_ref < run.in > run.log; e=$(date +%s.%N); echo "$ex plum_ref (CPU, 1 core) 2000 steps: $(echo "$e - $s" | bc) s" )
  fi
  rm -rf $d
done
