#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
for rep in 1 2; do
python tests/tools/perf_quick.py c2_500k
for v in b8 b6 m3 m0; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c2_500k; done
done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep3.log
