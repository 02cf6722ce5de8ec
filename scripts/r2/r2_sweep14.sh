#!/bin/bash
# round 2: peer-memory exchange (csrc/peer.cu) on one GPU (contexts of one process play the ranks) + no regression of the frame
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2d_pytest_peer.log 2>&1
tail -15 gpurun_out/r2d_pytest_peer.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/r2d_bench_g1_quick.json 2> gpurun_out/r2d_bench_g1_quick.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2d_bench_g1_quick.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "k_trace", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"])
P
