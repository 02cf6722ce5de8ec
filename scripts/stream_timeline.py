import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
from defectproj import Context, FrameStream, synth
K, H, W = synth.camera_wfov(); pose = synth.fill_frame_pose()
V, F = synth.param_mesh(*synth.MESH_CONFIGS["c2_500k"], seed=0, scale=6.0)
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
ctx.set_timing(True)
heats = [torch.ones((H, W)).pin_memory() for _ in range(4)]
N = 40
poses = np.stack([pose] * N)
fs = FrameStream(ctx, H, W, want=("pixel", "t_hit", "face"))
for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5): pass
fs.profile = True
for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5): pass
print("ms/frame", fs.last_elapsed_ms / N)
for i in range(20, 28):
    t = fs.timeline[i]
    print(i, " ".join(f"{x - fs.timeline[20][0]:7.3f}" for x in t), " | k dur %.3f  gap-to-prev-k %.3f" % (t[3] - t[2], t[2] - fs.timeline[i - 1][3]))
print("last_timings in stream mode:", ctx.last_timings())
hd = torch.ones((1, H, W), device="cuda"); n = H * W
o = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), t_hit=torch.empty(n, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"))
for _ in range(5): ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=True)
print("last_timings standalone:", ctx.last_timings())
# standalone + concurrent D2H traffic on another stream
s2 = torch.cuda.Stream()
a = torch.empty(n * 3, dtype=torch.int32, device="cuda"); h = torch.empty(n * 3, dtype=torch.int32).pin_memory()
for rep in range(3):
    with torch.cuda.stream(s2):
        for _ in range(4): h.copy_(a, non_blocking=True)
    ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=True)
    print("last_timings with concurrent D2H:", ctx.last_timings())
    torch.cuda.synchronize()
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
for rep in range(3):
    with torch.cuda.stream(s2):
        for _ in range(12): d2.copy_(h2, non_blocking=True)
    ctx.project_device(hd, K, pose[None], 0.5, "object", True, out=o, sync=True)
    print("last_timings with concurrent H2D:", ctx.last_timings())
    torch.cuda.synchronize()
