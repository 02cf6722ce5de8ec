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
    from defectproj.projector import combine_accumulators, fold_block, gather_hits, gather_slices, shard_range
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

        def frames_one(b):
            heat = synth.blob_heatmap((H, W), seed=b)
            xs, ys, I = orc.heatmap_to_points(heat, 0.4)
            return bvh.cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(K, poses[b])))

        lo, hi = shard_range(B, world, rank)
        (h, fm, vm), rec = frames(lo, hi)
        h, fm, vm = torch.from_numpy(h), torch.from_numpy(fm), torch.from_numpy(vm)
        combine_accumulators(h, fm, vm)
        allrec = gather_hits(torch.from_numpy(rec))
        (h1, f1, v1), rec1 = frames(0, B)
        ok = (np.array_equal(h.numpy(), h1) and np.array_equal(fm.numpy(), f1) and np.array_equal(vm.numpy(), v1)
              and np.array_equal(allrec.numpy(), rec1) and int(h1.sum()) == len(rec1) > 100)
        # the records to one rank only (the viewer's), unpadded; an empty contribution from the other rank
        only0 = gather_hits(torch.from_numpy(rec), dst=0)
        ok = ok and ((only0 is None) if rank != 0 else np.array_equal(only0.numpy(), rec1))
        part = torch.from_numpy(rec if rank == 0 else rec[:0])
        both = gather_hits(part)
        ok = ok and both.shape[0] == (len(rec) if rank == 0 else both.shape[0]) and both.shape[1] == 3

        # two batches, accumulated: every rank snapshots its block per batch (hist | fmax bits | vmax bits, the layout of
        # dp_accum_layout) and folds the combined snapshot into the totals -- each hit counted once (reducing
        # the live accumulators in place counted earlier batches `world` times)
        nF, nV = len(F), len(V)
        total = torch.zeros(2 * nF + nV, dtype=torch.int32)
        for b_lo, b_hi in ((0, 2), (2, B)):
            lo2, hi2 = shard_range(b_hi - b_lo, world, rank)
            (hb, fb, vb), _ = frames(b_lo + lo2, b_lo + hi2)
            snap = torch.from_numpy(np.concatenate([hb, fb.view(np.int32), vb.view(np.int32)]))
            fold_block(total, snap, nF)
        tn = total.numpy()
        ok = ok and (np.array_equal(tn[:nF], h1) and np.array_equal(tn[nF:2 * nF].view(np.float32), f1)
                     and np.array_equal(tn[2 * nF:].view(np.float32), v1))

        # single-frame ray sharding: each rank owns a slice of the frame's per-ray results, gathered in place
        full_t, full_f = frames_one(0)
        n = len(full_f)
        ranges = [shard_range(n, world, r) for r in range(world)]
        mine_f = torch.full((n,), -7, dtype=torch.int32)
        mine_f[ranges[rank][0]:ranges[rank][1]] = torch.from_numpy(full_f[ranges[rank][0]:ranges[rank][1]])
        gather_slices(mine_f, ranges)
        ok = ok and np.array_equal(mine_f.numpy(), full_f)
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
