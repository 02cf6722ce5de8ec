#!/bin/bash
# round 2, call 5: light-packets-last schedule, all gpu tests, the rewritten bench.py (N=1) + reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for m in c2_500k ns_1m c4_5m; do python tests/tools/perf_quick.py $m --check; done
} > gpurun_out/r2_sweep5.log 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_3.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_g1.json 2> gpurun_out/r2_bench_g1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_g1.err
cat gpurun_out/r2_sweep5.log; tail -15 gpurun_out/r2_pytest_gpu_3.log; tail -c 1500 gpurun_out/r2_bench_g1.err; head -c 6000 gpurun_out/r2_bench_g1.json; head -c 1500 gpurun_out/r2_bench_ref.json
