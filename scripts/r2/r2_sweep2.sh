#!/bin/bash
# round 2, call 2: uncompressed (208-byte) node twin vs the compressed set, same box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for rep in 1 2; do
for m in c2_500k ns_1m; do
  DP_FAT=0 python tests/tools/perf_quick.py $m --check
  DP_FAT=1 python tests/tools/perf_quick.py $m --check
done; done
python tests/tools/perf_quick.py c4_5m --check
python tests/tools/perf_quick.py c1_30k --check
} > gpurun_out/r2_sweep2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1
cat gpurun_out/r2_sweep2.log; tail -15 gpurun_out/r2_pytest_gpu_1.log
