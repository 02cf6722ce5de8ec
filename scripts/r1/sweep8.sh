#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
for v in b5 b4 b3; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c4_5m; done
python tests/tools/perf_quick.py ns_1m
for v in b5 b4; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py ns_1m; done
python tests/tools/perf_quick.py c2_500k
for v in b5; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c2_500k; done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep8.log
