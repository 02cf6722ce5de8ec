"""Warm build time (median of 9, the library's own CUDA events) per mesh; for A/B by environment variable."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth
for mesh in sys.argv[1:] or ["c1_30k", "c2_500k", "ns_1m", "c4_5m"]:
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0, scale=6.0)
    with Context(0) as ctx:
        ctx.set_mesh(V, F).build_bvh()
        ts = []
        for _ in range(9):
            ctx.build_bvh()
            ts.append(ctx.stats()["last_build_ms"])
        print(mesh, "DP_COOP_PER_SM", os.environ.get("DP_COOP_PER_SM", "default"), "build ms median %.4f min %.4f" % (np.median(ts), np.min(ts)), flush=True)
