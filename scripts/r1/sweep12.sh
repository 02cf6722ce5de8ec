#!/bin/bash
# tree rotations in k_binfit (DP_ROTATE_MAX = largest subtree that may rotate; 0 = off)
cd "$(dirname "$0")/../.."
{
for r in 0 64 512 8192 100000000; do
  echo "== DP_ROTATE_MAX=$r"
  DP_ROTATE_MAX=$r timeout 120 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
done
for p in 2 4; do
  echo "== DP_ROTATE_MAX=100000000 DP_ROTATE_PASSES=$p"
  DP_ROTATE_MAX=100000000 DP_ROTATE_PASSES=$p timeout 120 python tests/tools/perf_quick.py c2_500k 2>&1 | tail -1
done
echo "== DP_ROTATE_MAX=100000000 with parity check"
DP_ROTATE_MAX=100000000 timeout 200 python tests/tools/perf_quick.py c2_500k --check 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r1d_sweep_rotations.log
