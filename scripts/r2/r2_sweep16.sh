#!/bin/bash
# round 2: heavy packets as narrow items (DP_HEAVY_NARROW) -- parity suite, then A/B on the three meshes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2e_pytest_gpu.log
for hv in 0 1; do
  for mesh in c2_500k ns_1m c4_5m; do
    DP_HEAVY_NARROW=$hv timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu --no-configs --mesh $mesh > gpurun_out/r2e_bench_${mesh}_hv$hv.json 2> gpurun_out/r2e_bench_${mesh}_hv$hv.err
    python - $mesh $hv <<'P'
import json, sys
mesh, hv = sys.argv[1:3]
d = json.loads([l for l in open(f"gpurun_out/r2e_bench_{mesh}_hv{hv}.json") if l.startswith("{")][-1])
print(mesh, "heavy_narrow", hv, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "k_trace", round(d["roofline"]["kernel_ms"], 4), "e2e", round(d["e2e"]["value"]), "hist ok", d["checks"]["hist_total_equals_hits"])
P
  done
done
python scripts/shard_probe.py c2_500k c4_5m > gpurun_out/r2e_shard_probe.log 2>&1; cat gpurun_out/r2e_shard_probe.log | tail -32
