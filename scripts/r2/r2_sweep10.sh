#!/bin/bash
cd "$(dirname "$0")/../.."
{
for rep in 1 2 3; do
for m in c2_500k ns_1m; do
  python tests/tools/perf_quick.py $m
  DEFECTPROJ_LIB=$PWD/variants/libdp_fatgen.so python tests/tools/perf_quick.py $m
done; done
} > gpurun_out/r2_sweep10.log 2>&1
cat gpurun_out/r2_sweep10.log
