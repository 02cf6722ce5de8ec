#!/bin/bash
# round 2 closing call: smoke + bench (1 GPU) + reference arm + ncu launch list + one full capture of the traversal kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2n_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench_g1.json 2> gpurun_out/r2n_bench_g1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2>> gpurun_out/r2n_bench_g1.err
python bench.py --scaling strong --mesh c4_5m --steps 10 > gpurun_out/r2n_bench_strong_g1.json 2>> gpurun_out/r2n_bench_g1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_r2n_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 10 -c 2 -f -o gpurun_out/prof_r2n_trace \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_r2n_2.log 2>&1
tail -c 300 gpurun_out/r2_smoke.log; tail -c 600 gpurun_out/r2n_bench_g1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n_bench_g1.json").read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],4),"k_trace",round(d["roofline"]["kernel_ms"],4),"frac",round(d["roofline"]["frac"],3), round(d["roofline"]["frac_at_80B_nodes"],3),"e2e",{k:round(v["value"]) for k,v in d["e2e"]["modes"].items()}, "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["configs"].items(): print(k, json.dumps(v)[:400])
s=json.loads(open("gpurun_out/r2n_bench_strong_g1.json").read().strip().splitlines()[-1]); print("strong g1", s["value"], s["ms_per_step"], s["strong"].get("pipelined_ms_per_frame"))
PY
python scripts/shard_probe.py c2_500k c4_5m > gpurun_out/r2n_shard_probe.log 2>&1
python scripts/facade_latency.py > gpurun_out/r2n_facade_latency.log 2>&1; cp gpurun_out/facade_latency.json gpurun_out/r2n_facade_latency.json
grep '^{' gpurun_out/r2n_facade_latency.log | cut -c1-200
