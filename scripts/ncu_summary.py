#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS segments) of the traversal kernel. Usage: ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_active.avg','sm__cycles_active.max','sm__cycles_elapsed.avg','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    o = {"kernel": d.get("Kernel Name", "")[:50], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in keys:
        if k in d: o[k] = d[k] + " " + units[hdr.index(k)]
    out.append(o)
    for k, v in o.items(): print(f"{k:80s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; data = rows[2:]
ia, ie, it, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
seg = []
for i, r in enumerate(data):
    try: e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
    except Exception: continue
    p = r[ia].split(); op = p[1] if p[0].startswith('@') else p[0]
    if seg and seg[-1]['e'] == e:
        g = seg[-1]; g['n'] += 1; g['t'] += t; g['s'] += s; g['ops'].append(op); g['end'] = i
    else:
        seg.append(dict(e=e, n=1, t=t, s=s, ops=[op], start=i, end=i))
tot_e = sum(g['e'] * g['n'] for g in seg); tot_s = sum(g['s'] for g in seg) or 1
print("total warp inst", tot_e, "samples", tot_s)
segs = []
for g in seg:
    w = g['e'] * g['n']
    if w / tot_e > 0.006 or g['s'] / tot_s > 0.006:
        c = Counter(g['ops']).most_common(6)
        line = f"[{g['start']:4d}-{g['end']:4d}] n={g['n']:3d} exec={g['e']:9d} inst%={100*w/tot_e:5.1f} samples%={100*g['s']/tot_s:5.1f} avg_threads={g['t']/max(1,w):5.1f} {c}"
        print(line); segs.append(line)
if len(sys.argv) > 2:
    json.dump({"metrics": out, "sass_segments": segs}, open(sys.argv[2], "w"), indent=1)
