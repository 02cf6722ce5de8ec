"""Refit time (dp_pose_mesh: float64 posing + one-launch bottom-up fit, the library's own events) and build time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth
for mesh in sys.argv[1:] or ["c1_30k", "c2_500k", "ns_1m", "c4_5m"]:
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0)
    with Context(0) as ctx:
        ctx.set_mesh(V.astype(np.float64), F).build_bvh()
        poses = synth.fibonacci_poses(12, radius=600.0)
        ts, tb = [], []
        for i in range(12):
            ctx.pose_mesh(poses[i]); ctx.synchronize()
            ts.append(ctx.stats()["last_refit_ms"])
        for _ in range(5):
            ctx.build_bvh(); tb.append(ctx.stats()["last_build_ms"])
        print(mesh, "refit ms median %.4f min %.4f | build ms median %.4f" % (np.median(ts[2:]), np.min(ts[2:]), np.median(tb)), flush=True)
