#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "collapse or structure or morton or tiny or refit or full_size or duplicate" > gpurun_out/r2_pytest_build2.log 2>&1
{
for m in c1_30k c2_500k ns_1m c4_5m; do
  python tests/tools/perf_quick.py $m --check
done
DP_COLLAPSE_LAUNCHES=1 python tests/tools/perf_quick.py c2_500k
DP_COLLAPSE_LAUNCHES=1 python tests/tools/perf_quick.py c4_5m
} > gpurun_out/r2_sweep12.log 2>&1
tail -4 gpurun_out/r2_pytest_build2.log; cat gpurun_out/r2_sweep12.log
