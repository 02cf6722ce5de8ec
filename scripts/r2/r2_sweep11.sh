#!/bin/bash
cd "$(dirname "$0")/../.."
for m in c2_500k c4_5m; do
DP_COLLAPSE_LAUNCHES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_build_levels_$m.csv \
    python tests/tools/perf_quick.py $m > gpurun_out/ncu_levels_$m.log 2>&1
done
python - <<'PY'
import csv
for m in ("c2_500k","c4_5m"):
    rows=list(csv.reader(open(f'gpurun_out/r2_build_levels_{m}.csv')))
    for i,r in enumerate(rows):
        if 'Kernel Name' in r: h=r; start=i; break
    ki=h.index('Kernel Name'); mi=h.index('Metric Value'); gi=h.index('Grid Size')
    seq=[(r[ki].split('(')[0][-24:], float(r[mi].replace(',',''))/1e3, r[gi]) for r in rows[start+2:] if len(r)>mi and r[mi].replace(',','').replace('.','').isdigit()]
    idx=[i for i,(k,v,g) in enumerate(seq) if 'k_morton' in k]
    a=idx[2]; b=next(i for i in range(a,len(seq)) if 'k_fit_all' in seq[i][0])
    print(m)
    for k,v,g in seq[a:b+1]:
        if 'collapse' in k or 'binfit' in k or 'fit_all' in k or 'karras' in k: print(f"   {k:26s} {v:8.1f} us grid {g}")
PY
