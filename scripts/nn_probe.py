"""align_to_surface: ring search on the target grid against the tiled scan (DP_NN_GRID=0)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth  # noqa: E402

out = []
rng = np.random.default_rng(1)
with Context(0) as ctx:
    for nu, nv, nq in ((300, 200, 20000), (700, 500, 100000), (1000, 1000, 100000)):
        V, _ = synth.param_mesh(nu, nv, seed=9)
        tp = V.astype(np.float64)
        tn = rng.normal(size=tp.shape)
        q = tp[rng.integers(0, len(tp), nq)] + rng.normal(scale=0.5, size=(nq, 3))
        row = {"queries": nq, "target": len(tp)}
        res = {}
        for mode in ("1", "0"):
            os.environ["DP_NN_GRID"] = mode
            res[mode] = ctx.align_to_surface(q, tp, tn, 0.5)
            best = 1e30
            for _ in range(2):
                t0 = time.perf_counter()
                ctx.align_to_surface(q, tp, tn, 0.5)
                best = min(best, (time.perf_counter() - t0) * 1e3)
            row["grid_ms" if mode == "1" else "scan_ms"] = best
        row["identical"] = bool(all(np.array_equal(a, b) for a, b in zip(res["1"], res["0"])))
        out.append(row)
        print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nn_probe.json"), "w"), indent=1)
