#!/bin/bash
# Build one engine library per experiment patch (tools/experiments/*.patch) WITHOUT touching the tree:
#   tools/ab_build.sh whole_group_staging redux_extents "whole_group_staging+redux_extents"
# -> variants/<name>/libplum_b200.so (git-ignored, travels with gpurun).  "a+b" applies both patches in order.
# Then on a GPU:  gpurun --timeout 600 -- 'bash tools/ab_run.sh <tag> <name> ...'
set -e
REPO=$(cd "$(dirname "$0")/.." && pwd)
for spec in "$@"; do
  work=$(mktemp -d /tmp/ab_XXXXXX)
  mkdir -p "$work/plum_b200" "$work/include"
  cp -r "$REPO/plum_b200/csrc" "$work/plum_b200/csrc"
  cp "$REPO/include/plum_b200.h" "$work/include/"
  IFS='+' read -ra parts <<< "$spec"
  for p in "${parts[@]}"; do
    (cd "$work" && git apply --unsafe-paths -p1 --include='plum_b200/csrc/*' "$REPO/tools/experiments/$p.patch")
  done
  mkdir -p "$REPO/variants/$spec"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
       -o "$REPO/variants/$spec/libplum_b200.so" "$work/plum_b200/csrc/pg_engine.cu"
  echo "built variants/$spec/libplum_b200.so"
  rm -rf "$work"
done
