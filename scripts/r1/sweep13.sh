#!/bin/bash
# tree rotations, second sweep: grandchild <-> grandchild candidates (DP_ROTATE_GG), other mesh sizes
cd "$(dirname "$0")/../.."
R=100000000
{
for cfg in "0 0 1" "$R 0 1" "$R 1 1" "$R 1 2"; do
  set -- $cfg
  echo "== c2_500k DP_ROTATE_MAX=$1 DP_ROTATE_GG=$2 DP_ROTATE_PASSES=$3"
  DP_ROTATE_MAX=$1 DP_ROTATE_GG=$2 DP_ROTATE_PASSES=$3 timeout 120 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
done
for mesh in ns_1m c4_5m; do
for cfg in "0 0 1" "$R 0 1" "$R 1 1"; do
  set -- $cfg
  echo "== $mesh DP_ROTATE_MAX=$1 DP_ROTATE_GG=$2 DP_ROTATE_PASSES=$3"
  DP_ROTATE_MAX=$1 DP_ROTATE_GG=$2 DP_ROTATE_PASSES=$3 timeout 150 python tests/tools/perf_quick.py $mesh 2>&1 | tail -1
done; done
echo "== c2_500k gg parity check"
DP_ROTATE_MAX=$R DP_ROTATE_GG=1 timeout 200 python tests/tools/perf_quick.py c2_500k --check 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r1d_sweep_rotations2.log
