#!/bin/bash
# round 2: A/B of the fused prologue (DP_FUSED_PROLOGUE) and the look-back width (DP_LB_WIDE: two builds), same box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2f_pytest_gpu.log
for lib in libdefectproj.so libdefectproj_lb1.so; do
  for fp in 1 0; do
    for rep in 1 2; do
      DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib DP_FUSED_PROLOGUE=$fp timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > gpurun_out/r2f_ab.json 2> gpurun_out/r2f_ab.err
      python - $lib $fp <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2f_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], "fused", sys.argv[2], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5))
P
    done
  done
done
