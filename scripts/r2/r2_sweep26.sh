#!/bin/bash
# round 2: prefetch, when a group is pushed, of the node its pop will visit (DP_PUSH_PF = 0 / 1 list 0 / 2 lists 0+1 / 3 all)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in libdefectproj.so libdefectproj_pp1.so libdefectproj_pp2.so libdefectproj_pp3.so libdefectproj.so libdefectproj_pp1.so; do
  for mesh in c2_500k ns_1m; do
    DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-configs --mesh $mesh > gpurun_out/r2q_ab.json 2> gpurun_out/r2q_ab.err
    python - $lib $mesh <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2q_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], sys.argv[2], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5), d["clocks"])
P
  done
done
for lib in libdefectproj.so libdefectproj_pp1.so libdefectproj_pp3.so; do
  echo "== shard probe $lib"; DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib python scripts/shard_probe.py c2_500k 2>&1 | grep -E '"world": (1|8), "rank": (0|2|5)'
done
