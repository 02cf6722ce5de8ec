"""Traversal time vs ray count for sparse frames: run with DP_NARROW=1 and DP_NARROW=0 to place the switch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
K, H, W = synth.camera_720p(); pose = synth.fixed_pose(z=350.0)
V, F = synth.param_mesh(*synth.MESH_CONFIGS["c2_500k"], seed=0)
ctx = Context(0); ctx.set_timing(True); ctx.set_mesh(V, F).build_bvh()
n = H * W
o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
out = []
for sigma in (30, 45, 60, 75, 90, 110, 130, 160):
    heat = torch.from_numpy(synth.gaussian_heatmap((H, W), sigma=float(sigma), dtype=np.float32))[None].cuda()
    ts = []
    for _ in range(8):
        nr, nh = ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=o, sync=True)
        ts.append(ctx.last_timings()["trace_ms"])
    out.append((nr, nh, round(1e3 * float(np.median(ts[2:])), 1)))
print("narrow", os.environ.get("DP_NARROW", "1"), out, flush=True)
