#!/bin/bash
cd "$(dirname "$0")/../.."
{
for m in 0 2 3 4 1; do DP_COLLAPSE=$m python tests/tools/perf_quick.py c2_500k --check; done
for m in 0 2 3 4 1; do DP_COLLAPSE=$m python tests/tools/perf_quick.py c4_5m; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep2.log
