#!/bin/bash
# round 2, call 8: in-kernel ray generation + hit points (DP_FUSE_RAYS A/B through bench.py), all gpu tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_4.log 2>&1
for f in 0 1; do
DP_FUSE_RAYS=$f python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/r2_bench_fuse$f.json 2> gpurun_out/r2_bench_fuse$f.err
done
python scripts/facade_latency.py > gpurun_out/r2b_facade_latency.log 2>&1; cp gpurun_out/facade_latency.json gpurun_out/r2b_facade_latency.json
python tests/tools/configs_report.py > gpurun_out/r2_configs_report.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu_4.log
python - <<'PY'
import json
for f in (0,1):
    d=json.loads(open(f"gpurun_out/r2_bench_fuse{f}.json").read().strip().splitlines()[-1])
    print("fuse",f,"value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"k_trace",round(d["roofline"]["kernel_ms"],4),"e2e",{k:round(v["value"]) for k,v in d["e2e"]["modes"].items()})
PY
grep '^{' gpurun_out/r2b_facade_latency.log
tail -5 gpurun_out/r2_configs_report.log
