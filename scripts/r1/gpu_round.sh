#!/bin/bash
# smoke + bench (1 GPU) + reference arm + ncu launch list + one full capture of the traversal kernel
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r1d.log 2> gpurun_out/bench_r1d.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1d.log 2>> gpurun_out/bench_r1d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1d_bench_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 10 -c 2 -f -o gpurun_out/prof_r1d_trace \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -c 600 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench_r1d.log; tail -c 1200 gpurun_out/bench_ref_r1d.log
