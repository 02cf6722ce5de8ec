#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
python tests/tools/perf_quick.py c2_500k --check
for rep in 1 2; do
python tests/tools/perf_quick.py c2_500k
DEFECTPROJ_LIB=$L/libdefectproj_nopf.so python tests/tools/perf_quick.py c2_500k
done
python tests/tools/perf_quick.py c4_5m
DEFECTPROJ_LIB=$L/libdefectproj_nopf.so python tests/tools/perf_quick.py c4_5m
} 2>&1 | grep -v Warning | tee gpurun_out/sweep4.log
