#!/bin/bash
# round 2: CTAs per SM of the uncompressed-node traversal (DP_MIN_BLOCKS_FAT = 4 / 5 / 6) now that the launch is known to
# follow its heaviest packets (fewer co-resident warps = a faster chain, no spills at 96+ registers)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in libdefectproj.so libdefectproj_mb5.so libdefectproj_mb4.so libdefectproj.so libdefectproj_mb5.so; do
  for mesh in c2_500k ns_1m; do
    DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$lib timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-configs --mesh $mesh > gpurun_out/r2p_ab.json 2> gpurun_out/r2p_ab.err
    python - $lib $mesh <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2p_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], sys.argv[2], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5))
P
  done
done
