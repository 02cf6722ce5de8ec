"""CPU, world_size 2, gloo: the host-side multi-GPU logic (frame sharding + result combination) with the
per-rank results produced by the oracle in place of a GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
    import torch
    import torch.distributed as dist
    from defectproj import synth
    from defectproj.projector import combine_accumulators, gather_hits, shard_range
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, F = synth.param_mesh(24, 16, seed=3)
        K, H, W = synth.K_matrix(150.0, 150.0, 80.0, 60.0), 120, 160
        B = 5
        poses = synth.fibonacci_poses(B, radius=400.0)
        bvh = orc.Bvh(V, F)

        def frames(lo, hi):
            fs, Is, recs = [], [], []
            for b in range(lo, hi):
                heat = synth.blob_heatmap((H, W), seed=b)
                xs, ys, I = orc.heatmap_to_points(heat, 0.4)
                t, f = bvh.cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(K, poses[b])))
                hit = f >= 0
                fs.append(f); Is.append(I)
                recs.append(np.stack([(b * H * W + ys * W + xs)[hit].astype(np.float64), f[hit].astype(np.float64),
                                      t[hit].astype(np.float64)], axis=1))
            f_all = np.concatenate(fs) if fs else np.zeros(0, np.int32)
            I_all = np.concatenate(Is) if Is else np.zeros(0, np.float32)
            rec = np.concatenate(recs) if recs else np.zeros((0, 3))
            return orc.accumulate(f_all, I_all, F, len(V)), rec

        lo, hi = shard_range(B, world, rank)
        (h, fm, vm), rec = frames(lo, hi)
        h, fm, vm = torch.from_numpy(h), torch.from_numpy(fm), torch.from_numpy(vm)
        combine_accumulators(h, fm, vm)
        allrec = gather_hits(torch.from_numpy(rec))
        (h1, f1, v1), rec1 = frames(0, B)
        ok = (np.array_equal(h.numpy(), h1) and np.array_equal(fm.numpy(), f1) and np.array_equal(vm.numpy(), v1)
              and np.array_equal(allrec.numpy(), rec1) and int(h1.sum()) == len(rec1) > 100)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_combination_equals_single_rank(orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
