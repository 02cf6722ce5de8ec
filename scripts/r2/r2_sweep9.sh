#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_5.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_5.log
