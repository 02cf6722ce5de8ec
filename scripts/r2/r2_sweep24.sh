#!/bin/bash
# round 2: VERY heavy packets (longest ray > DP_NARROW_FRAC x mean) as eight narrow items in a pre-pass; A/B against
# DP_HEAVY_NARROW=0 and thresholds 3 / 4 / 6, three meshes, same box; then the shard probe
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2o_pytest_gpu.log
run() {  # lib hv mesh
  DEFECTPROJ_LIB=$PWD/6dof-pose-estimation-and-defect-projection_b200/defectproj/$1 DP_HEAVY_NARROW=$2 timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-configs --mesh $3 > gpurun_out/r2o_ab.json 2> gpurun_out/r2o_ab.err
  python - $1 $2 $3 <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2o_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], "heavy_narrow", sys.argv[2], sys.argv[3], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5), "hist ok", d["checks"]["hist_total_equals_hits"])
P
}
for mesh in c2_500k ns_1m c4_5m; do
  run libdefectproj.so 0 $mesh; run libdefectproj.so 1 $mesh; run libdefectproj_nf3.so 1 $mesh; run libdefectproj_nf6.so 1 $mesh; run libdefectproj.so 0 $mesh; run libdefectproj.so 1 $mesh
done
python scripts/shard_probe.py c2_500k c4_5m 2>&1 | grep -E '"world": (1|2|8),' > gpurun_out/r2o_shard_probe.log; cat gpurun_out/r2o_shard_probe.log
