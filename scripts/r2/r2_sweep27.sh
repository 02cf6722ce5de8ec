#!/bin/bash
# round 2: software-pipelined eight-lane path (next node fetched before the current node's triangle tests)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2r_pytest_gpu.log
for lib in libdefectproj_base.so libdefectproj.so libdefectproj_base.so libdefectproj.so; do
  echo "== $lib"; DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib python scripts/sparse_probe.py 2>&1 | tail -4 | cut -c1-250
done
