#!/bin/bash
# with the rotations on: hybrid-collapse threshold and the rotation floor once more
cd "$(dirname "$0")/../.."
{
for h in 128 2048 8192; do
  echo "== c2_500k DP_HYBRID_COUNT=$h"
  DP_HYBRID_COUNT=$h timeout 150 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
done
for mn in 256 1024; do
  echo "== c2_500k DP_ROTATE_MIN=$mn"
  DP_ROTATE_MIN=$mn timeout 150 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r1d_sweep_rotations4.log
