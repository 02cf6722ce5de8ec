#!/bin/bash
# tree rotations, third sweep: only subtrees of at least DP_ROTATE_MIN triangles (the top of the tree)
cd "$(dirname "$0")/../.."
R=100000000
{
for mesh in c2_500k c4_5m; do
for mn in 64 1024 16384; do
  echo "== $mesh DP_ROTATE_MIN=$mn"
  DP_ROTATE_MAX=$R DP_ROTATE_MIN=$mn timeout 150 python tests/tools/perf_quick.py $mesh 2>&1 | tail -1
done; done
echo "== c2_500k DP_ROTATE_MIN=1024 passes 3"
DP_ROTATE_MAX=$R DP_ROTATE_MIN=1024 DP_ROTATE_PASSES=3 timeout 150 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r1d_sweep_rotations3.log
