#!/bin/bash
# round 2: FrameStream read-back queued on a prediction (no host round trip between kernels and D2H)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frame_stream" 2>&1 | tail -3
for rep in 1 2; do
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs > gpurun_out/r2g_bench_g1_quick.json 2> gpurun_out/r2g_bench_g1_quick.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2g_bench_g1_quick.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "e2e", round(d["e2e"]["value"]), "pcie", round(d["e2e"]["pcie_d2h_gbs"], 1), "bound", round(d["e2e"]["pcie_bound_mrays_s"]), {k: (round(v["value"]), round(v["ms_per_frame"], 4)) for k, v in d["e2e"]["modes"].items()})
P
done
