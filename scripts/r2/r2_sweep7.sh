#!/bin/bash
# round 2, call 7: traversal micro-variants (atomic queue slots, explicit shared loads of the culling bound, L1 prefetch of the next node)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for rep in 1 2; do
for m in c2_500k ns_1m; do
  python tests/tools/perf_quick.py $m
  for v in tqa basm fpf all3; do
    DEFECTPROJ_LIB=$PWD/variants/libdp_$v.so python tests/tools/perf_quick.py $m --check
  done
done; done
} > gpurun_out/r2_sweep7.log 2>&1
cat gpurun_out/r2_sweep7.log
