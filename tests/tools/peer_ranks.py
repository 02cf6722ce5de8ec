"""Real ranks over CUDA IPC (run under torchrun on a box with >= 2 GPUs; tests/test_gpu_peer.py launches it):
the peer-memory combine and the peer-stored ray-sharded frame against a local context that does all the work alone.
NCCL is used for the 64-byte handle exchange and for this script's own pass/fail vote only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))


def frame_pose(i):
    from defectproj import synth
    return synth.look_at_pose(eye=(6 * 60.0 + 7.0 * np.sin(0.7 * i), -30.0 + 15 * np.cos(0.3 * i), 6 * 8.0), target=(6 * 30.0, 6 * 52.0, 3.0 * i))


def main():
    import torch
    import torch.distributed as dist
    from defectproj import Context, Projector, synth
    from defectproj.projector import BatchCombiner
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    mesh = os.environ.get("PEER_MESH", "c1_30k")
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[mesh], seed=0, scale=6.0)
    H, W = 512, 1024
    K = synth.K_matrix(126.0 * W / 512, 126.0 * W / 512, W / 2, H / 2)
    n_px = H * W
    ok = True
    msgs = []

    def check(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            msgs.append(f"rank {rank}: FAILED {what}")

    proj = Projector(V, F, device=local)
    check(proj.peer is not None, "Projector did not enable the peer path: " + getattr(proj, "peer_error", ""))
    if proj.peer is not None:
        check(proj.enable_peer(record_rows=n_px, result_rays=n_px), "enable_peer with records + results")
    if proj.peer is None:
        print("\n".join(msgs), flush=True)
        dist.destroy_process_group()
        sys.exit(1)
    ctx, comb = proj.ctx, proj.combiner
    ref = Context(local).set_mesh(V, F).build_bvh()
    heat = torch.rand((1, H, W), device=dev)
    dist.broadcast(heat, 0)
    out = dict(pixel=torch.empty(n_px, dtype=torch.int32, device=dev), t_hit=torch.empty(n_px, device=dev),
               face=torch.empty(n_px, dtype=torch.int32, device=dev))
    gathered = torch.zeros((world * n_px, 3), dtype=torch.int32, device=dev) if rank == 0 else None
    total_rows = torch.zeros(1, dtype=torch.int64).pin_memory()

    # ---- A: batches of frames, combined by ONE kernel per rank
    ref.accum_reset()
    ctx.accum_reset()
    comb.reset_totals()
    for batch in range(3):
        want_rows = []
        for r in range(world):
            nfr = 1 + (r + batch) % 2
            for j in range(nfr):
                pose = frame_pose(10 * batch + 3 * r + j)[None]
                if r == rank:
                    n, h = ctx.project_device(heat, K, pose, 0.5, "object", True, out=out, sync=True)
                ro = dict(pixel=torch.empty(n_px, dtype=torch.int32, device=dev), t_hit=torch.empty(n_px, device=dev),
                          face=torch.empty(n_px, dtype=torch.int32, device=dev))
                nr, hr = ref.project_device(heat, K, pose, 0.5, "object", True, out=ro, sync=True)
            want_rows.append(ref.pack_records_device(ro["t_hit"], ro["face"], pixel=ro["pixel"], n=nr).clone())
        k = comb.acquire()
        rec, cnt = comb.records(k)
        ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n, out=rec, count_async=cnt, sync=False)
        comb.submit(gather_root=0, gathered=gathered, count_async=total_rows)
        a, b, c = proj.combined()
        torch.cuda.synchronize()
        comb.check()
        hist, fmax, vmax = ref.accum_get()
        check(np.array_equal(a.cpu().numpy(), hist), f"batch {batch}: histogram totals")
        check(np.array_equal(b.cpu().numpy(), fmax) and np.array_equal(c.cpu().numpy(), vmax), f"batch {batch}: maxima")
        if rank == 0:
            want = torch.cat(want_rows, dim=0)
            check(int(total_rows[0]) == want.shape[0], f"batch {batch}: gathered row count {int(total_rows[0])} vs {want.shape[0]}")
            check(torch.equal(gathered[:want.shape[0]], want), f"batch {batch}: gathered records")

    # ---- A2: the same through the public call: Projector.project_batch(records_to=0) / hit_records()
    for use_peer in (True, False):
        if not use_peer:
            proj.peer.ctx.peer_close()                       # collective (every rank is here): back to the NCCL combiner
            dist.barrier()
            proj.peer = None
            proj.combiner = BatchCombiner(ctx, None)
        B = 2
        heats = torch.rand((B, H, W), device=dev) * (0.6 + 0.1 * rank)
        poses_b = np.stack([frame_pose(50 + 2 * rank + j) for j in range(B)])
        ob = dict(pixel=torch.empty(B * n_px, dtype=torch.int32, device=dev), t_hit=torch.empty(B * n_px, device=dev),
                  face=torch.empty(B * n_px, dtype=torch.int32, device=dev))
        if use_peer:
            check(proj.enable_peer(record_rows=B * n_px, result_rays=n_px), "enable_peer for a batch of two frames")
        nb, hb = proj.project_batch(heats, K, poses_b, 0.5, out=ob, records_to=0)
        mine = ctx.pack_records_device(ob["t_hit"], ob["face"], pixel=ob["pixel"], n=nb)
        check(mine.shape[0] == hb, "records of this rank == its hit count")
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = mine.shape[0]
        dist.all_reduce(sizes)
        allrec = torch.zeros((int(sizes.sum()), 3), dtype=torch.int32, device=dev)
        lo = int(sizes[:rank].sum())
        allrec[lo:lo + mine.shape[0]] = mine
        dist.all_reduce(allrec)                              # disjoint rows: the sum is the concatenation in rank order
        got = proj.hit_records()
        hist_b = proj.combined()[0]
        torch.cuda.synchronize()
        if rank == 0:
            check(got is not None and got.shape == allrec.shape and torch.equal(got, allrec), f"hit_records() (peer={use_peer})")
        else:
            check(got is None, "hit_records() is None off the root")
        tot = torch.tensor([hb], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        check(int(hist_b.sum()) == int(tot[0]), f"combined histogram counts every rank's hits (peer={use_peer})")
    check(proj.enable_peer(record_rows=n_px, result_rays=n_px), "peer path back on")
    ctx, comb = proj.ctx, proj.combiner

    # ---- B: ONE frame's rays sharded, results stored into every rank's window by the traversal
    for case in ("dense", "sparse"):
        hm = torch.ones((H, W), device=dev) if case == "dense" else torch.from_numpy(synth.blob_heatmap((H, W), seed=5)).to(dev)
        for frame in range(3):
            pose = frame_pose(frame)
            whole = dict(t_hit=torch.empty(n_px, device=dev), face=torch.empty(n_px, dtype=torch.int32, device=dev),
                         point=torch.empty((n_px, 3), device=dev))
            ref.accum_reset()
            n, h = ref.project_device(hm[None], K, pose[None], 0.5, "object", True, out=whole, sync=True)
            res = {"point": None}
            nn, hh, _ = proj.project_frame_sharded(hm, K, pose, 0.5, out=res, gather="peer", reduce=True, reset=True)
            hist = proj.combined()[0]
            torch.cuda.synchronize()
            comb.check()
            check(int(nn) == n, f"{case} frame {frame}: ray count")
            check(torch.equal(res["face"][:n], whole["face"][:n]), f"{case} frame {frame}: faces")
            check(torch.equal(res["t_hit"][:n].view(torch.int32), whole["t_hit"][:n].view(torch.int32)), f"{case} frame {frame}: t_hit")
            check(torch.equal(res["point"][:n].view(torch.int32), whole["point"][:n].view(torch.int32)), f"{case} frame {frame}: points")
            check(np.array_equal(hist.cpu().numpy(), ref.accum_get()[0]) and h > 100, f"{case} frame {frame}: histogram")

    # ---- timing, for the record: the same combine through NCCL and through peer memory
    nccl = BatchCombiner(ctx, None)
    stream = torch.cuda.current_stream()

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def via_nccl():
        nccl.submit(stream, reset=False)
        nccl.result(stream)

    def via_peer():
        comb.submit(stream, reset=False)
        comb.result(stream)

    t_nccl, t_peer = timed(via_nccl), timed(via_peer)
    comb.check()
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for m in msgs:
        print(m, flush=True)
    if rank == 0:
        print(f"combine of the accumulator block ({comb.words * 4} bytes, {world} ranks): NCCL {t_nccl:.3f} ms, peer kernel {t_peer:.3f} ms", flush=True)
        print("PEER_RANKS_OK" if int(flag[0]) == 1 else "PEER_RANKS_FAILED", flush=True)
    ref.close()
    proj.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
