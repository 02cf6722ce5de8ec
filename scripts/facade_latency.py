"""Wall time of the drop-in call itself -- defect_projection.ray_tracing(data_dir, mesh, heatmap, intrinsics, thr) with
the reference's own argument types (float64 720p heatmap, float64 vertices, threshold 0.75 as in run.py:115) -- next to
the CPU port of the same call (per-call BVH build, as the reference does).  cProfile of one call shows where the host
side spends it."""
import cProfile
import io
import json
import os
import pstats
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import defect_projection as dpj, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out = []
K, H, W = synth.camera_720p()
heat = synth.gaussian_heatmap((H, W), dtype=np.float64)
pose = synth.fixed_pose()
with tempfile.TemporaryDirectory() as d:
    c2d = np.eye(4)
    c2d[:3, 3] = [32.0, 2.0, -4.0]
    synth.write_scene_dir(d, K, (H, W), color_to_depth=c2d)
    for mesh in ("c1_30k", "c2_500k"):
        V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0)
        Vd = orc.transform_points(V.astype(np.float64), c2d @ pose)          # mesh in the depth-camera frame (run.py:109-110)
        tm = dpj.TriangleMesh(Vd, F)
        for thr in (0.75, 0.5):
            pcd, _ = dpj.ray_tracing(d, tm, heat, K, thr)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                pcd, _ = dpj.ray_tracing(d, tm, heat, K, thr)
                ts.append((time.perf_counter() - t0) * 1e3)
            # CPU port of the same call: pose + per-call BVH build + frame
            t0 = time.perf_counter()
            Vc = orc.pose_vertices(Vd, np.linalg.inv(c2d))
            r = orc.Bvh(Vc, F).project_frame(heat.astype(np.float32), thr, K)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            row = {"mesh": mesh, "threshold": thr, "rays": int(dpj.last_result()["n_rays"]), "hits": len(pcd.points),
                   "ray_tracing_ms_median": float(np.median(ts)), "ray_tracing_ms_min": float(min(ts)),
                   "cpu_port_ms": cpu_ms, "cpu_cores": orc.num_threads()}
            out.append(row)
            print(json.dumps(row), flush=True)
        # the heatmap already on the GPU (HeatmapReader.get_heatmap(device=True) prepares it there from the 224x224 raw map)
        import torch
        heat_d = torch.from_numpy(heat).cuda()
        for thr in (0.75, 0.5):
            pcd, _ = dpj.ray_tracing(d, tm, heat_d, K, thr)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                pcd, _ = dpj.ray_tracing(d, tm, heat_d, K, thr)
                ts.append((time.perf_counter() - t0) * 1e3)
            row = {"mesh": mesh, "threshold": thr, "heatmap": "device-resident", "rays": int(dpj.last_result()["n_rays"]),
                   "hits": len(pcd.points), "ray_tracing_ms_median": float(np.median(ts)), "ray_tracing_ms_min": float(min(ts))}
            out.append(row)
            print(json.dumps(row), flush=True)
        # the production pattern: the same model at a new pose on every call (run.py:109-110) -> refit, not rebuild
        moved = []
        for k in range(12):
            Tk = np.eye(4)
            Tk[:3, :3] = synth.rot_z(3.0 * k)
            moved.append(dpj.TriangleMesh(orc.transform_points(Vd, Tk), F))
        dpj.ray_tracing(d, moved[0], heat, K, 0.75)
        ts = []
        for m in moved[1:]:
            t0 = time.perf_counter()
            dpj.ray_tracing(d, m, heat, K, 0.75)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts_rebuild = []
        for m in moved[1:]:
            dpj._SCENE["F"] = None                                   # what the call cost when every new pose meant a rebuild
            t0 = time.perf_counter()
            dpj.ray_tracing(d, m, heat, K, 0.75)
            ts_rebuild.append((time.perf_counter() - t0) * 1e3)
        row = {"mesh": mesh, "threshold": 0.75, "moving_mesh_refit_ms_median": float(np.median(ts)),
               "moving_mesh_rebuild_ms_median": float(np.median(ts_rebuild))}
        out.append(row)
        print(json.dumps(row), flush=True)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(5):
            dpj.ray_tracing(d, tm, heat, K, 0.75)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
        print(s.getvalue()[:3500], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "facade_latency.json"), "w"), indent=1)
