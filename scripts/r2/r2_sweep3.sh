#!/bin/bash
# round 2, call 3: occupancy / generic-visit variants of the uncompressed-node traversal, the new full-size float64
# tests, and one full ncu capture of the new kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for m in c2_500k ns_1m; do
  python tests/tools/perf_quick.py $m
  for v in fat6 fat8 fatgen; do
    DEFECTPROJ_LIB=$PWD/variants/libdp_$v.so python tests/tools/perf_quick.py $m
  done
done
} > gpurun_out/r2_sweep3.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_northstar.py tests/test_open3d_pin.py -m gpu -x -q -rs > gpurun_out/r2_pytest_northstar.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/prof_r2a_trace \
    python tests/tools/perf_quick.py c2_500k > gpurun_out/ncu_r2a.log 2>&1
cat gpurun_out/r2_sweep3.log; tail -15 gpurun_out/r2_pytest_northstar.log; tail -3 gpurun_out/ncu_r2a.log
