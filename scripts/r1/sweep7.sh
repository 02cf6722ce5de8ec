#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj
{
for rep in 1 2; do
python tests/tools/perf_quick.py c4_5m
for v in pfc b8 b8pfc b5; do DEFECTPROJ_LIB=$L/libdefectproj_$v.so python tests/tools/perf_quick.py c4_5m; done
done
} 2>&1 | grep -v Warning | tee gpurun_out/sweep7.log
