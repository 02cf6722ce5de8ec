#!/bin/bash
# knob sweep on the GPU box: collapse policy, triangle/node cost ratio, triangle-queue flush threshold
cd "$(dirname "$0")/../.."
L=6dof-pose-estimation-and-defect-projection_b200/defectproj
{
DP_COLLAPSE=0 python tests/tools/perf_quick.py c2_500k --check
for c in 0.2 0.35 0.5 0.75 1.0 1.5; do DP_COLLAPSE=1 DP_CPRIM=$c python tests/tools/perf_quick.py c2_500k; done
DP_COLLAPSE=1 DP_CPRIM=0.5 python tests/tools/perf_quick.py c2_500k --check
for v in tq8 tq16 tq24; do DEFECTPROJ_LIB=$PWD/$L/libdefectproj_$v.so python tests/tools/perf_quick.py c2_500k; done
DP_COLLAPSE=0 python tests/tools/perf_quick.py ns_1m
DP_COLLAPSE=1 python tests/tools/perf_quick.py ns_1m
DP_COLLAPSE=0 python tests/tools/perf_quick.py c4_5m
DP_COLLAPSE=1 python tests/tools/perf_quick.py c4_5m
} 2>&1 | grep -v Warning | tee gpurun_out/sweep1.log
