"""Timeline of FrameStream on the bench workload (per frame: H2D begin/end, kernels begin/end, D2H begin/end, ms since
frame 20's H2D began), with the L2 flush of bench.py on the kernel stream, for the payload and the accumulate-only mode."""
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
N = 40
poses = np.stack([bench.frame_pose(i) for i in range(N)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fl(_s): flush[:132 << 20].zero_()
for want in (("pixel", "face", "point"), ()):
    for use_flush in (True, False):
        fs = FrameStream(ctx, H, W, want=want)
        for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5, before_kernels=fl if use_flush else None): pass
        fs.profile = True
        for i, r in fs.run([heats[i % 4] for i in range(N)], K, poses, 0.5, before_kernels=fl if use_flush else None): pass
        print("want", want, "flush", use_flush, "ms/frame", round(fs.last_elapsed_ms / N, 4))
        for i in range(20, 26):
            t = fs.timeline[i]
            print(i, " ".join(f"{x - fs.timeline[20][0]:7.3f}" for x in t), " | k %.3f  k-gap %.3f  d2h %.3f" % (t[3] - t[2], t[2] - fs.timeline[i - 1][3], t[5] - t[4]))
