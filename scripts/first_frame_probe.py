"""Traversal time of an isolated frame (no learnt schedule: DP_ORDER=0) with and without the strided packet walk."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
K, H, W = synth.camera_wfov(); pose = synth.fill_frame_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2_500k"], seed=0, scale=6.0)
ctx = Context(0); ctx.set_timing(True); ctx.set_mesh(V, F).build_bvh()
heat = torch.ones((1, H, W), dtype=torch.float32, device="cuda"); n = H * W
o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(8):
    flush.zero_()
    ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=o, sync=True)
    ts.append(ctx.last_timings()["trace_ms"])
print("order", os.environ.get("DP_ORDER", "1"), "spread", os.environ.get("DP_SPREAD", "1"), "trace_ms", [round(t, 4) for t in ts], flush=True)
