"""Stage timings of one sparse frame (the reference's production case: 720p Gaussian heatmap, threshold 0.5 / 0.75)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, synth
K, H, W = synth.camera_720p(); pose = synth.fixed_pose()
for mesh in ("c1_30k", "c2_500k"):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0)
    ctx = Context(0); ctx.set_timing(True); ctx.set_mesh(V, F).build_bvh()
    heat = torch.from_numpy(synth.gaussian_heatmap((H, W), dtype=np.float32))[None].cuda()
    n = H * W
    o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
    for thr in (0.5, 0.75):
        for _ in range(5): nr, nh = ctx.project_device(heat, K, pose[None], thr, "object", True, out=o, sync=True)
        tl = ctx.last_timings()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): ctx.project_device(heat, K, pose[None], thr, "object", True, out=o, sync=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
        t0 = time.perf_counter()
        for _ in range(200): ctx.project_device(heat, K, pose[None], thr, "object", True, out=o, sync=True)
        dts = (time.perf_counter() - t0) / 200
        print(mesh, "thr", thr, "rays", nr, "hits", nh, {k: round(v * 1e3, 1) for k, v in tl.items()}, "us | back-to-back %.1f us/frame, with sync %.1f us/frame" % (dt * 1e6, dts * 1e6), flush=True)
