#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
python tests/tools/perf_quick.py c2_500k
for v in s6 s4 s6q128 s14; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c2_500k; done
python tests/tools/perf_quick.py c2_500k
python tests/tools/perf_quick.py c4_5m
for v in s6 s6q128; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c4_5m; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep11.log
