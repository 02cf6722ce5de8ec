#!/bin/bash
cd "$(dirname "$0")/../.."
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/nccl_probe.py > gpurun_out/r2_nccl_probe_g$N.log 2>&1
grep -v "^\[\|Warning\|warn\|\*\*\*\|OMP_NUM" gpurun_out/r2_nccl_probe_g$N.log | tail -14
bash scripts/r2/r2_multi.sh $N
