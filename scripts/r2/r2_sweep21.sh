#!/bin/bash
# round 2: per-column / per-row tables of the ray generation's two pixel quotients (DP_RAYTAB), A/B on one box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2i_pytest_gpu.log
for rt in 1 0 1 0; do
  DP_RAYTAB=$rt timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > gpurun_out/r2i_ab.json 2> gpurun_out/r2i_ab.err
  python - $rt <<'P'
import json, sys
d = json.loads([l for l in open("gpurun_out/r2i_ab.json") if l.startswith("{")][-1])
print("raytab", sys.argv[1], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "k_trace", round(d["roofline"]["kernel_ms"], 5), "e2e", round(d["e2e"]["value"]))
P
done
