#!/bin/bash
# round 2, 2 GPUs: the peer-memory exchange between real ranks (CUDA IPC), then the bench line at 2 GPUs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/tools/peer_ranks.py > gpurun_out/r2d_peer_ranks_g$N.log 2>&1
echo "peer_ranks rc=$?"; grep -v '^W\|^\*\*\*\|OMP_NUM' gpurun_out/r2d_peer_ranks_g$N.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2d_bench_g$N.json 2> gpurun_out/r2d_bench_g$N.err
echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench_g$N.err
python - $N <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2d_bench_g{n}.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "ms/step", d["ms_per_step"], "e2e", {k: round(v["value"]) for k, v in d["e2e"]["modes"].items()})
    print("combine", {k: v for k, v in d["combine"].items() if k != "what"})
    for k, v in d.get("configs", {}).items():
        print(k, {a: b for a, b in v.items() if a not in ("timed",)})
except Exception as e:
    print("no line:", e)
P
