#!/bin/bash
cd "$(dirname "$0")/../.."
{
nvidia-smi topo -m 2>&1 | head -20
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:40])"
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist) $(cat $d/class); fi; done | head -20
lscpu | grep -i "numa\|socket\|model name" | head
} > gpurun_out/numa_probe.log 2>&1
for b in 1 0; do
DP_NUMA_BIND=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$b bench.py --gpus 8 --steps 60 --warmup 5 2>/dev/null > gpurun_out/bench_g8_numa$b.log
done
