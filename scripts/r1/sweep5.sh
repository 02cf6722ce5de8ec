#!/bin/bash
cd "$(dirname "$0")/../.."
{
for h in 0 24 64 192 512 4096; do DP_HYBRID_COUNT=$h python tests/tools/perf_quick.py c2_500k --check; done
for h in 0 64 512; do DP_HYBRID_COUNT=$h python tests/tools/perf_quick.py c4_5m; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep5.log
