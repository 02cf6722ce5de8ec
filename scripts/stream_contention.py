"""Does the copy traffic slow the kernels of a pipelined frame?  FrameStream on the bench workload without the L2
flush; per variant the mean duration of a frame's kernels (events around dp_project on the kernel stream):
payload outputs with their read-back / the same outputs written but never read back / no H2D either (one resident
heatmap) / accumulate-only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch
import bench
from defectproj import Context, FrameStream
V, F, K, H, W = bench.workload("c2_500k")
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
heats = [torch.ones((H, W)).pin_memory() for _ in range(4)]
N = 60
poses = np.stack([bench.frame_pose(i) for i in range(N)])
def run(want, d2h=True, h2d=True):
    fs = FrameStream(ctx, H, W, want=want)
    if not d2h:
        for hbuf in fs.out_h:
            for k in list(hbuf):
                hbuf[k] = hbuf[k][:0]            # zero-length pinned views: the copies move nothing
        fs.cap_copy = 0
    for rep in range(2):
        fs.profile = rep == 1
        for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5): pass
    k = [t[3] - t[2] for t in fs.timeline[10:]]
    return fs.last_elapsed_ms / N, float(np.mean(k)), float(np.min(k))
print("payload, read back      ms/frame %.4f  kernels mean %.4f min %.4f" % run(("pixel", "face", "point")))
print("accumulate-only         ms/frame %.4f  kernels mean %.4f min %.4f" % run(()))
# outputs written, never read back: the device-resident loop on the same stream discipline
stream = torch.cuda.current_stream()
n = H * W
out = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), face=torch.empty(n, dtype=torch.int32, device="cuda"), point=torch.empty((n, 3), device="cuda"))
heat = torch.ones((1, H, W), device="cuda")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
for rep in range(2):
    for i in range(N):
        ev[i][0].record(stream); ctx.project_device(heat, K, poses[i][None], 0.5, "object", True, out=out); ev[i][1].record(stream)
    torch.cuda.synchronize()
k = [a.elapsed_time(b) for a, b in ev[10:]]
print("payload outputs, no copies at all: kernels mean %.4f min %.4f" % (np.mean(k), np.min(k)))
# the same loop with an unrelated D2H / H2D stream running
s2 = torch.cuda.Stream()
big_d = torch.empty(16 << 20, dtype=torch.uint8, device="cuda"); big_h = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
for what in ("d2h", "h2d"):
    for rep in range(2):
        with torch.cuda.stream(s2):
            for _ in range(80):
                (big_h.copy_(big_d, non_blocking=True) if what == "d2h" else big_d.copy_(big_h, non_blocking=True))
        for i in range(N):
            ev[i][0].record(stream); ctx.project_device(heat, K, poses[i][None], 0.5, "object", True, out=out); ev[i][1].record(stream)
        torch.cuda.synchronize()
    k = [a.elapsed_time(b) for a, b in ev[10:]]
    print("payload outputs, unrelated %s copies of 16 MB back to back: kernels mean %.4f min %.4f" % (what, np.mean(k), np.min(k)))
