#!/bin/bash
# round 2, call 1: today's baseline (GPU tests + traversal timing) and the octant-specialised visit, back to back
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for rep in 1 2; do
for m in c2_500k ns_1m c4_5m; do
  python tests/tools/perf_quick.py $m --check
  DEFECTPROJ_LIB=$PWD/variants/libdp_oct.so python tests/tools/perf_quick.py $m --check
done; done
} > gpurun_out/r2_sweep1.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_0.log 2>&1
tail -30 gpurun_out/r2_sweep1.log; tail -5 gpurun_out/r2_pytest_gpu_0.log
