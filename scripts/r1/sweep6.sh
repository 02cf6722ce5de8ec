#!/bin/bash
cd "$(dirname "$0")/../.."
{
for c in 0.25 0.35 0.5 0.7 1.0 1.4; do DP_CPRIM=$c python tests/tools/perf_quick.py c2_500k; done
for h in 128 2048 32768; do DP_HYBRID_COUNT=$h python tests/tools/perf_quick.py c2_500k; done
for c in 0.35 0.5 0.7; do DP_CPRIM=$c python tests/tools/perf_quick.py ns_1m; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep6.log
