#!/usr/bin/env python
"""One-line traversal timing of configs[1] for knob sweeps (DP_REFILL, DP_TILED, DEFECTPROJ_LIB)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
mesh = sys.argv[1] if len(sys.argv) > 1 else "c2_500k"
check = "--check" in sys.argv
K, H, W = synth.camera_wfov()
pose = synth.fill_frame_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0, scale=6.0)
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
ctx.set_timing(True)
builds = []
for _ in range(4):
    ctx.build_bvh(); builds.append(ctx.stats()["last_build_ms"])
heat = torch.ones((1, H, W), dtype=torch.float32, device="cuda")
n = H * W
out = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(8):
    flush.zero_()
    ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
    ts.append(ctx.last_timings()["trace_ms"])
ts = ts[2:]
ctx.set_stats(True)
ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
st = ctx.stats()
ctx.set_stats(False)
msg = f"{mesh} hybrid={os.environ.get('DP_HYBRID_COUNT','-')} collapse={os.environ.get('DP_COLLAPSE','-')} cprim={os.environ.get('DP_CPRIM','-')} wide_nodes={st['n_wide_nodes']} depth={st['wide_depth']} build_ms={min(builds):.2f} refill={os.environ.get('DP_REFILL','-')} tiled={os.environ.get('DP_TILED','-')} lib={os.path.basename(os.environ.get('DEFECTPROJ_LIB','default'))} trace_ms min {min(ts):.4f} med {np.median(ts):.4f} -> {n/np.median(ts)/1e3:.0f} Mrays/s nodes/ray {st['nodes_fetched']/max(1,st['rays']):.2f} tris/ray {st['tris_tested']/max(1,st['rays']):.2f}"
if check:
    from oracle import oracle as orc
    xs = np.tile(np.arange(W, dtype=np.int64), H); ys = np.repeat(np.arange(H, dtype=np.int64), W)
    t, f = orc.Bvh(V, F).cast_f32(orc.rays_object_frame(xs, ys, Context.frame_xform(K, pose)))
    msg += f" parity face={np.array_equal(out['face'].cpu().numpy(), f)} t={np.array_equal(out['t_hit'].cpu().numpy().view(np.uint32), t.view(np.uint32))}"
print(msg, flush=True)
