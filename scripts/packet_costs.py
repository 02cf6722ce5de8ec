#!/usr/bin/env python
"""Per-ray node counts of the benchmark frame (counting kernel variant) -> gpurun_out/ray_nodes_<mesh>.npy, for the
CPU-side schedule simulation (scripts/schedule_sim.py)."""
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
c = ctx.ray_node_counts(n)
np.save(os.path.join(ROOT, "gpurun_out", f"ray_nodes_{mesh}.npy"), c.astype(np.uint8 if c.max() < 256 else np.uint16))
print(mesh, "mean", c.mean(), "max", c.max())
