#!/bin/bash
# round 2, call 6: cooperative single-launch collapse (A/B vs per-level launches), build launch lists
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "collapse or structure or morton or tiny or refit or full_size" > gpurun_out/r2_pytest_build.log 2>&1
{
for m in c1_30k c2_500k ns_1m c4_5m; do
  DP_COLLAPSE_LAUNCHES=1 python tests/tools/perf_quick.py $m
  DP_COLLAPSE_LAUNCHES=0 python tests/tools/perf_quick.py $m --check
done
} > gpurun_out/r2_sweep6.log 2>&1
for m in c2_500k c4_5m; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_build_launches_$m.csv \
    python tests/tools/perf_quick.py $m > gpurun_out/ncu_build_$m.log 2>&1
done
python scripts/facade_latency.py > gpurun_out/r2_facade_latency.log 2>&1; cp gpurun_out/facade_latency.json gpurun_out/r2_facade_latency.json
timeout 600 python -m pytest tests -m gpu -x -q -k 'facade or tracker or refits' > gpurun_out/r2_pytest_facade.log 2>&1
tail -5 gpurun_out/r2_pytest_build.log; tail -5 gpurun_out/r2_pytest_facade.log; grep -v '^ ' gpurun_out/r2_facade_latency.log | grep '^{'; cat gpurun_out/r2_sweep6.log
