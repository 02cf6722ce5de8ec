#!/bin/bash
# round 2: compaction tile size (DP_CT_CHUNKS = 4 / 8 / 16: 4096 / 8192 / 16384 pixels per tile), three builds, same box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in libdefectproj.so libdefectproj_ct8.so libdefectproj_ct16.so; do
  export DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact or project_object or batch_of_frames" 2>&1 | tail -1
  for rep in 1 2; do
    timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > gpurun_out/r2f_ab.json 2> gpurun_out/r2f_ab.err
    python - $lib <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2f_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5), "c1 e2e", round(d["e2e"]["value"]))
P
  done
done
