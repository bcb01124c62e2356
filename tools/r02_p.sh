#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cp plum_b200/libplum_b200.so /tmp/base.so
for v in k8t384b2 k8t448b2 k8t512b2 k8t320b3; do
  cp variants/$v/libplum_b200.so plum_b200/libplum_b200.so
  echo "== $v"; timeout 600 python tools/chain_probe.py --system S --steps 1000 --clusters "" --replicas 296,444 2>&1 | tee gpurun_out/r02p_$v.jsonl
done
cp /tmp/base.so plum_b200/libplum_b200.so
