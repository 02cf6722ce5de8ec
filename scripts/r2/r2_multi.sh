#!/bin/bash
# multi-GPU bench: N from $1 (default 2), weak + strong lines
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/r2_bench_ref_g$N.json 2>> gpurun_out/r2_bench_g$N.err
tail -c 1500 gpurun_out/r2_bench_g$N.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_g$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", {k:round(v["value"]) for k,v in d["e2e"]["modes"].items()})
print("combine", {k:v for k,v in d["combine"].items() if k!="what"})
print("checks", d["checks"])
print("configs", json.dumps(d.get("configs"), indent=0)[:1500])
r=json.loads(open("gpurun_out/r2_bench_ref_g$N.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
