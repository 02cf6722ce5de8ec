#!/bin/bash
# round 2, call 4: all gpu tests on the new code (sharding, records, combiner, float64 truth), leaf cost sweep, packet costs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_2.log 2>&1
{
for m in c2_500k ns_1m; do
  for cp in 0.5 0.75 1.0 1.5; do
    echo "DP_CPRIM=$cp"; DP_CPRIM=$cp python tests/tools/perf_quick.py $m
  done
done
python scripts/packet_costs.py c2_500k
} > gpurun_out/r2_sweep4.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_2.log; cat gpurun_out/r2_sweep4.log
