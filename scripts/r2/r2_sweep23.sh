#!/bin/bash
# round 2: host heatmaps go through a chunked pinned upload and the device-resident path of the drop-in call
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prepost.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
python scripts/facade_latency.py > gpurun_out/r2k_facade_latency.log 2>&1; cp gpurun_out/facade_latency.json gpurun_out/r2k_facade_latency.json
grep '^{' gpurun_out/r2k_facade_latency.log | cut -c1-260
DP_FACADE_UPLOAD=0 python scripts/facade_latency.py 2>&1 | grep '^{' | grep -v device-resident | cut -c1-200
