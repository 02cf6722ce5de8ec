"""One GPU: how the frame time scales with the shard size (dp_set_ray_shard(rank, world)) -- the kernels of one rank of a
ray-sharded frame without any exchange.  Prints per (mesh, world, rank): ms per frame (CUDA events around dp_project, L2
flushed before each frame) and the traversal kernel's own time."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from defectproj import Context  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
rows = []
for mesh in sys.argv[1:] or ["c2_500k", "c4_5m"]:
    V, F, K, H, W = bench.workload(mesh)
    n_pix = H * W
    with Context(0) as ctx:
        ctx.set_mesh(V, F).build_bvh()
        heat = torch.ones((1, H, W), device="cuda")
        out = dict(t_hit=torch.empty(n_pix, device="cuda"), face=torch.empty(n_pix, dtype=torch.int32, device="cuda"))
        poses = [bench.frame_pose(i) for i in range(16)]
        for world in (1, 2, 4, 8):
            for rank in range(world):
                ctx.set_ray_shard(rank, world)
                ms, kms = [], []
                for i in range(13):
                    flush.zero_()
                    timing = i >= 3 and i % 3 == 0
                    ctx.set_timing(timing)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    ctx.project_device(heat, K, poses[i][None], 0.5, "object", True, out=out, sync=False)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    if timing:
                        kms.append(ctx.last_timings()["trace_ms"])
                    elif i >= 3:
                        ms.append(e0.elapsed_time(e1))
                ctx.set_timing(False)
                row = {"mesh": mesh, "world": world, "rank": rank, "ms_per_frame": round(float(np.mean(ms)), 4),
                       "trace_ms": round(float(np.mean(kms)), 4)}
                rows.append(row)
                print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "shard_probe.json"), "w"), indent=1)
