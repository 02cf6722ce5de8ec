import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
mesh = sys.argv[1] if len(sys.argv) > 1 else "c2_500k"
K, H, W = synth.camera_wfov()
pose = synth.fill_frame_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0, scale=6.0)
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
heat = torch.ones((1, H, W), dtype=torch.float32, device="cuda")
n = H * W
out = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
ctx.set_stats(True)
ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
c = ctx.ray_node_counts(n).reshape(H, W)
print("nodes/ray mean", c.mean(), "median", np.median(c), "p90", np.percentile(c, 90), "p99", np.percentile(c, 99), "p99.9", np.percentile(c, 99.9), "max", c.max())
print("hist:", np.histogram(c, bins=[0, 5, 10, 15, 20, 30, 50, 100, 200, 500, 100000])[0])
# warp-level: max over 32x1 strips and 8x4 tiles
rows = c.reshape(H, W // 32, 32).max(2)
tiles = c.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
print("lockstep efficiency rows 32x1:", c.sum() / (rows.sum() * 32), " tiles 8x4:", c.sum() / (tiles.sum() * 32))
print("max per strip: mean", rows.mean(), "max", rows.max(), " per tile: mean", tiles.mean(), "max", tiles.max())
ys, xs = np.nonzero(c > np.percentile(c, 99.9))
print("heavy rays bbox: y", ys.min(), ys.max(), "x", xs.min(), xs.max())
# per-block-of-16-rows cost
print("cost by 64-row band:", c.reshape(16, 64, W).sum(axis=(1, 2)) / c.sum())
t = out["t_hit"].cpu().numpy().reshape(H, W)
print("t range", t.min(), t.max())
