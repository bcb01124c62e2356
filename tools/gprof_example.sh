#!/bin/bash
# flat gprof profile of the host driver (bin/plum_gpu_prof, built with PLUM_HOST_GPROF=1) on one example
ex=${1:-bulk_nvt}
d=$(mktemp -d); cp tests/golden/examples/$ex/* $d/; sed -i 's/^s1_total_simulation_steps .*/s1_total_simulation_steps 100000/' $d/run.in
( cd $d; PLUM_SEED=1 $GRAFT_REPO_ROOT/bin/plum_gpu_prof < run.in > run.log; gprof -b -p $GRAFT_REPO_ROOT/bin/plum_gpu_prof gmon.out | head -${2:-25} )
rm -rf $d
