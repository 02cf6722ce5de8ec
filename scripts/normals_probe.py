"""Normal estimation (dp_estimate_normals) timing from host buffers, with the reference's three parameter sets."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth  # noqa: E402

out = []
with Context(0) as ctx:
    for nu, nv in ((300, 200), (500, 500), (1000, 1000)):
        V, _ = synth.param_mesh(nu, nv, seed=9)
        P = V.astype(np.float64)
        for radius, max_nn in ((10.0, 30), (2.0, 5), (0.1, 30)):
            _, cnt = ctx.estimate_normals(P, radius, max_nn, want_counts=True)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                ctx.estimate_normals(P, radius, max_nn)
                best = min(best, (time.perf_counter() - t0) * 1e3)
            row = {"points": len(P), "radius": radius, "max_nn": max_nn, "ms": best, "mean_neighbours": float(cnt.mean())}
            out.append(row)
            print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "normals_probe.json"), "w"), indent=1)
