#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
for rep in 1 2; do
python tests/tools/perf_quick.py c2_500k
DEFECTPROJ_LIB=$L/libdefectproj_b6.so python tests/tools/perf_quick.py c2_500k
done
python tests/tools/perf_quick.py ns_1m
DEFECTPROJ_LIB=$L/libdefectproj_b6.so python tests/tools/perf_quick.py ns_1m
} 2>&1 | grep -v Warning | tee gpurun_out/sweep10.log
