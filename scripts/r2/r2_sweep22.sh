#!/bin/bash
# round 2: heavy packets (list 0 / lists 0+1 of the learnt schedule) prefetch their hit children and triangles towards L1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in libdefectproj_hp0.so libdefectproj.so libdefectproj_hp2.so libdefectproj_hp0.so libdefectproj.so; do
  export DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib
  for mesh in c2_500k c4_5m; do
    timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-configs --mesh $mesh > gpurun_out/r2j_ab.json 2> gpurun_out/r2j_ab.err
    python - $lib $mesh <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2j_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], sys.argv[2], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5))
P
  done
done
for lib in libdefectproj_hp0.so libdefectproj.so libdefectproj_hp2.so; do
  export DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib
  echo "== shard probe $lib"; python scripts/shard_probe.py c2_500k c4_5m 2>&1 | grep -E '"world": (1|8), "rank": (0|2|5)'
done
