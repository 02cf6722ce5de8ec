#!/usr/bin/env python
"""Measures the BASELINE.json configs other than the headline one (bench.py covers configs[1]) on one GPU
and writes gpurun_out/configs_report.json.  Every GPU result is checked against the oracle where the oracle
finishes in seconds."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
from oracle import oracle as orc

out = {}
ctx = Context(0)
ctx.set_timing(True)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

# ---------------------------------------------------------------- C1
K, H, W = synth.camera_720p(); pose = synth.fixed_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
ctx.set_mesh(V, F).build_bvh()
c1 = {"triangles": len(F), "build_ms": ctx.stats()["last_build_ms"]}
for name, heat in (("gaussian", synth.gaussian_heatmap((H, W), dtype=np.float32)), ("dense", synth.dense_heatmap((H, W)))):
    hd = torch.from_numpy(heat)[None].cuda()
    n = H * W
    o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
    nr, nh = ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=True)
    ms = timed(lambda: ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=False))
    t0 = time.perf_counter(); bv = orc.Bvh(orc.pose_vertices(V.astype(np.float64), pose), F); r = bv.project_frame(heat, 0.5, K); cpu = time.perf_counter() - t0
    xs, ys, I = orc.heatmap_to_points(heat, 0.5)
    t, f = orc.Bvh(V, F).cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(K, pose)))
    c1[name] = {"rays": nr, "hits": nh, "gpu_ms_per_frame": ms, "gpu_mrays_s": nr / ms / 1e3, "cpu_ms_per_frame_incl_build": cpu * 1e3,
                "cpu_cores": orc.num_threads(), "parity_face": bool(np.array_equal(o["face"][:nr].cpu().numpy(), f))}
out["C1"] = c1
print("C1", json.dumps(c1), flush=True)

# ---------------------------------------------------------------- C3: 64 views, refit per frame
V, F = synth.param_mesh(*synth.MESH_CONFIGS["c2_500k"], seed=0)
ctx.set_mesh(V, F).build_bvh()
B = 64
poses = synth.fibonacci_poses(B, radius=600.0)
heats = torch.from_numpy(np.stack([synth.blob_heatmap((H, W), seed=i) for i in range(B)])).cuda()
def c3_object():
    ctx.accum_reset(torch.cuda.current_stream())
    ctx.project_device(heats, K, poses, 0.5, "object", True, out=None, sync=False)
def c3_camera():
    ctx.accum_reset(torch.cuda.current_stream())
    for b in range(B):
        ctx.pose_mesh(poses[b], torch.cuda.current_stream())
        ctx.project_device(heats[b:b + 1], K, None, 0.5, "camera", True, out=None, sync=False)
nr, nh = ctx.project_device(heats, K, poses, 0.5, "object", True, out=None, sync=True)
ms_o = timed(c3_object, reps=5, warm=1)
c3_object(); torch.cuda.synchronize(); h_obj = ctx.accum_get()[0].copy()
ms_c = timed(c3_camera, reps=3, warm=1)
c3_camera(); torch.cuda.synchronize(); h_cam = ctx.accum_get()[0].copy()
ctx.pose_mesh(poses[0]); torch.cuda.synchronize(); refit_ms = ctx.stats()["last_refit_ms"]
out["C3"] = {"views": B, "triangles": len(F), "rays_total": nr, "hits_total": nh,
             "object_frame_single_launch_ms_per_frame": ms_o / B, "object_mrays_s": nr / ms_o / 1e3,
             "camera_frame_refit_per_frame_ms_per_frame": ms_c / B, "camera_mrays_s": nr / ms_c / 1e3, "refit_ms": refit_ms,
             "hist_sum_object": int(h_obj.sum()), "hist_sum_camera": int(h_cam.sum()),
             "hist_faces_differing_object_vs_camera": int((h_obj != h_cam).sum())}
print("C3", json.dumps(out["C3"]), flush=True)

# ---------------------------------------------------------------- C4: 5M triangles, build + traversal
Kw, Hw, Ww = synth.camera_wfov(); posew = synth.fill_frame_pose()
for cfg in ("ns_1m", "c4_5m"):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[cfg], seed=0, scale=6.0)
    Vd, Fd = torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda()
    ctx.set_mesh(Vd, Fd)
    bms = []
    for _ in range(4):
        ctx.build_bvh(); bms.append(ctx.stats()["last_build_ms"])
    st = ctx.stats()
    hd = torch.ones((1, Hw, Ww), device="cuda"); n = Hw * Ww
    o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
    ctx.set_stats(True); nr, nh = ctx.project_device(hd, Kw, posew[None], 0.5, "object", True, out=o, sync=True); s2 = ctx.stats(); ctx.set_stats(False)
    tms = []
    for _ in range(8):
        ctx.project_device(hd, Kw, posew[None], 0.5, "object", True, out=o, sync=True); tms.append(ctx.last_timings()["trace_ms"])
    xs = np.tile(np.arange(Ww, dtype=np.int64), Hw); ys = np.repeat(np.arange(Hw, dtype=np.int64), Ww)
    t0 = time.perf_counter(); bv = orc.Bvh(V, F); t1 = time.perf_counter(); t, f = bv.cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(Kw, posew))); t2 = time.perf_counter()
    out["C4_" + cfg] = {"triangles": len(F), "build_ms_first": bms[0], "build_ms": float(np.median(bms[1:])), "mtris_s": len(F) / np.median(bms[1:]) / 1e3,
                        "wide_nodes": st["n_wide_nodes"], "depth": st["wide_depth"], "bvh_bytes": st["n_wide_nodes"] * 80 + len(F) * 48,
                        "rays": nr, "hit_frac": nh / nr, "trace_ms": float(np.median(tms[2:])), "mrays_s": nr / np.median(tms[2:]) / 1e3,
                        "nodes_per_ray": s2["nodes_fetched"] / s2["rays"], "tris_per_ray": s2["tris_tested"] / s2["rays"],
                        "cpu_build_s": t1 - t0, "cpu_cast_s": t2 - t1, "cpu_cores": orc.num_threads(),
                        "parity_face": bool(np.array_equal(o["face"].cpu().numpy(), f)),
                        "parity_t": bool(np.array_equal(o["t_hit"].cpu().numpy().view(np.uint32), t.view(np.uint32)))}
    print(cfg, json.dumps(out["C4_" + cfg]), flush=True)
    del bv
# ---------------------------------------------------------------- the rows either side of the path (8f), host-buffer calls
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
rng = np.random.default_rng(0)
fr = {}
data = rng.random((224, 224))
fr["prepare_heatmap_224_to_1024x1024_f64_ms"] = wall(lambda: ctx.prepare_heatmap(data, 1024, 1024))
d_data = torch.from_numpy(data).cuda()
fr["prepare_heatmap_device_ms"] = wall(lambda: ctx.prepare_heatmap(d_data, 1024, 1024, np.float32))
t0 = time.perf_counter(); orc.prepare_heatmap(data, 1024, 1024); fr["prepare_heatmap_cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
n = 1 << 20
I = rng.random(n).astype(np.float32); face = rng.integers(-1, 1000, n).astype(np.int32); p64 = rng.normal(size=(n, 3))
fr["pack_hits_1M_host_buffers_ms"] = wall(lambda: ctx.pack_hits(I, face, None, p64, T=np.eye(4)), reps=3)
t0 = time.perf_counter(); orc.pack_hits(I, face, p64, np.eye(4)); fr["pack_hits_cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
Vi, Fi = synth.param_mesh(300, 200, seed=9)
Vi = Vi.astype(np.float64)
fn_ = np.cross(Vi[Fi[:, 1]] - Vi[Fi[:, 0]], Vi[Fi[:, 2]] - Vi[Fi[:, 0]]); vn = np.zeros_like(Vi)
for k in range(3): np.add.at(vn, Fi[:, k], fn_)
vn /= np.linalg.norm(vn, axis=1, keepdims=True)
Ti = np.eye(4); Ti[:3, :3] = synth.rot_z(1.5) @ synth.rot_x(-1.0); Ti[:3, 3] = [0.4, -0.3, 0.5]
src = (Vi[::3] @ np.linalg.inv(Ti)[:3, :3].T + np.linalg.inv(Ti)[:3, 3])
r = ctx.icp_point_to_plane(src, Vi, vn, 3.0)
ms = wall(lambda: ctx.icp_point_to_plane(src, Vi, vn, 3.0), reps=2)
fr["icp_20k_x_60k"] = {"ms_total": ms, "iterations": r["iterations"], "ms_per_iteration": ms / (r["iterations"] + 1),
                       "fitness": r["fitness"], "inlier_rmse": r["inlier_rmse"], "max_abs_error_vs_applied_pose": float(np.abs(r["transformation"] - Ti).max())}
fr["estimate_normals_60k_radius2_nn5_ms"] = wall(lambda: ctx.estimate_normals(Vi, 2.0, 5), reps=3)
fr["estimate_normals_60k_radius10_nn30_ms"] = wall(lambda: ctx.estimate_normals(Vi, 10.0, 30), reps=3)
q_ = Vi[rng.integers(0, len(Vi), 20000)] + rng.normal(scale=0.5, size=(20000, 3))
fr["align_to_surface_20k_x_60k_ms"] = wall(lambda: ctx.align_to_surface(q_, Vi, vn, 0.5), reps=3)
out["f_rows"] = fr
print("f_rows", json.dumps(fr), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_report.json"), "w"), indent=1)
