import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, FrameStream, synth
K, H, W = synth.camera_wfov(); pose = synth.fill_frame_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS["c2_500k"], seed=0, scale=6.0)
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
heats = [torch.ones((H, W)).pin_memory() for _ in range(4)]
N = 60
poses = np.stack([pose] * N)
for want in (("face",), ("t_hit", "face"), ("pixel", "t_hit", "face")):
    fs = FrameStream(ctx, H, W, want=want)
    for rep in range(2):
        t0 = time.perf_counter()
        for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5): pass
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(want, "ms/frame events", fs.last_elapsed_ms / N, "wall", 1e3 * wall / N, flush=True)
# host-side cost of one asynchronous dp_project (no sync)
hd = torch.ones((1, H, W), device="cuda"); n = H * W
o = dict(t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
for _ in range(3): ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=False)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("enqueue host ms/call", 1e3 * (t1 - t0) / 50, "total ms/call", 1e3 * (t2 - t0) / 50)
# raw copy bandwidth
a = torch.empty(n * 3, dtype=torch.int32, device="cuda"); h = torch.empty(n * 3, dtype=torch.int32).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): h.copy_(a, non_blocking=True)
torch.cuda.synchronize(); print("D2H 12.6MB ms", 1e3 * (time.perf_counter() - t0) / 20)
for on in (True, False, True, False):
    ctx.set_timing(on)
    for _ in range(5): ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=False)
    torch.cuda.synchronize(); print("stage events", on, "-> %.4f ms/call" % (1e3 * (time.perf_counter() - t0) / 200), flush=True)
ctx.set_timing(True)
