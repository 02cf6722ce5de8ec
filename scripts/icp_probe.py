"""ICP row: uniform-grid nearest-neighbour search against the tiled scan of the whole target
(DP_ICP_GRID=0), same call, same data, back to back.  Prints one JSON line per size to gpurun_out/icp_probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth  # noqa: E402


def cloud(nu, nv, seed):
    V, F = synth.param_mesh(nu, nv, seed=seed)
    V = V.astype(np.float64)
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    return V, vn


def wall(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


out = []
T = np.eye(4)
T[:3, :3] = synth.rot_z(1.5) @ synth.rot_x(-1.0)
T[:3, 3] = [0.4, -0.3, 0.5]
Ti = np.linalg.inv(T)
with Context(0) as ctx:
    for nu, nv, step, dist in ((40, 25, 1, 3.0), (300, 200, 3, 3.0), (300, 200, 3, 0.75), (700, 500, 2, 3.0), (1000, 1000, 2, 1.0)):
        V, vn = cloud(nu, nv, seed=9)
        src = V[::step] @ Ti[:3, :3].T + Ti[:3, 3]
        row = {"source": len(src), "target": len(V), "max_correspondence_distance": dist}
        res = {}
        for mode in ("1", "0"):
            if mode == "0" and len(V) * len(src) > 2.5e10:
                continue                                            # the scan of 1M x 500k points takes seconds per iteration
            os.environ["DP_ICP_GRID"] = mode
            r = ctx.icp_point_to_plane(src, V, vn, dist, want_correspondence=True)
            ms = wall(lambda: ctx.icp_point_to_plane(src, V, vn, dist))
            res[mode] = r
            key = "grid" if mode == "1" else "scan"
            row[key + "_ms_total"] = ms
            row[key + "_ms_per_evaluation"] = ms / (r["iterations"] + 1)
            row["iterations"] = r["iterations"]
            row["fitness"] = r["fitness"]
            row["pose_error"] = float(np.abs(r["transformation"] - T).max())
        if len(res) == 2:
            row["identical"] = bool(np.array_equal(res["1"]["correspondence"], res["0"]["correspondence"]) and
                                    np.array_equal(res["1"]["transformation"], res["0"]["transformation"]))
        out.append(row)
        print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "icp_probe.json"), "w"), indent=1)
