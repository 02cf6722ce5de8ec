#!/bin/bash
cd "$(dirname "$0")/../.."
{
for rep in 1 2; do
for m in 0 1; do
  echo "== DP_COLLAPSE=$m rep $rep"
  DP_COLLAPSE=$m python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'kernel_ms',round(r['kernel_ms'],4),'nodes',round(r['nodes_per_ray'],2),'tris',round(r['tris_per_ray'],2),'frac',round(r['frac'],3),'build',round(d['bvh']['build_ms'],2))"
done; done
} 2>&1 | tee gpurun_out/ab_collapse.log
