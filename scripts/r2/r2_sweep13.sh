#!/bin/bash
# round 2: device-resident heatmap through the drop-in call
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prepost.py -m gpu -x -q > gpurun_out/r2_pytest_prepost.log 2>&1
python scripts/facade_latency.py > gpurun_out/r2c_facade_latency.log 2>&1; cp gpurun_out/facade_latency.json gpurun_out/r2c_facade_latency.json
tail -12 gpurun_out/r2_pytest_prepost.log; grep '^{' gpurun_out/r2c_facade_latency.log
