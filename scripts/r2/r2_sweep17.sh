#!/bin/bash
# round 2: compaction that does dp_project's resets itself (no prologue launch) + 128-wide look-back rounds
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/r2f_bench_g1_quick.json 2> gpurun_out/r2f_bench_g1_quick.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2f_bench_g1_quick.json") if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "k_trace", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], "pcie", d["e2e"].get("pcie_d2h_gbs"), {k: round(v["value"]) for k, v in d["e2e"]["modes"].items()})
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2f_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_r2f_1.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2f_bench_launches.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:60]].append(float(r[iv].replace(",", "")))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:12]:
    print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.2f} us")
P
