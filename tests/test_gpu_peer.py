"""Result exchange over peer-mapped memory (csrc/peer.cu, include/defectproj.h "dp_peer_*").

On ONE GPU several contexts of this process play the ranks (dp_peer_open_local maps their windows by address; the
kernels, flags and epochs are the ones the multi-process path runs).  With >= 2 GPUs the same checks run as real ranks
under torchrun (CUDA IPC handles, NCCL only for the handle exchange): tests/tools/peer_ranks.py.
The combination must be bit-identical to the 1-GPU result: integer sums and float maxima are order-independent."""
import os
import subprocess
import sys

import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene():
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0, scale=6.0)
    H, W = 256, 512
    K = synth.K_matrix(126.0 * W / 512, 126.0 * W / 512, W / 2, H / 2)
    return V, F, H, W, K


def _frame_pose(i):
    return synth.look_at_pose(eye=(6 * 60.0 + 7.0 * np.sin(0.7 * i), -30.0 + 15 * np.cos(0.3 * i), 6 * 8.0), target=(6 * 30.0, 6 * 52.0, 3.0 * i))


@pytest.mark.parametrize("world", [2, 3])
def test_peer_combine_equals_one_gpu(built_lib, world, monkeypatch):
    """Every rank projects its own frames; per batch ONE dp_peer_combine per rank folds all snapshots into every rank's
    totals and gathers the hit records on rank 0: totals == the single context that projected every frame, records ==
    the ranks' records concatenated in rank order; two batches (both slots, totals keep accumulating)."""
    import torch
    from defectproj import Context
    from defectproj.projector import PeerCombiner
    monkeypatch.setenv("DP_PEER_BLOCKS", "16")          # all ranks' grids co-resident on the one device
    V, F, H, W, K = _scene()
    n_px = H * W
    heat = torch.rand((1, H, W), device="cuda")
    ctxs = [Context(0).set_mesh(V, F).build_bvh() for _ in range(world)]
    ref = Context(0).set_mesh(V, F).build_bvh()
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        combs = [PeerCombiner(c, r, world, record_rows=n_px, local_contexts=ctxs) for r, c in enumerate(ctxs)]
        for cb in combs:
            cb.open_local()
        outs = [dict(pixel=torch.empty(n_px, dtype=torch.int32, device="cuda"), t_hit=torch.empty(n_px, device="cuda"),
                     face=torch.empty(n_px, dtype=torch.int32, device="cuda")) for _ in range(world)]
        gathered = torch.zeros((world * n_px, 3), dtype=torch.int32, device="cuda")
        total_rows = torch.zeros(1, dtype=torch.int64).pin_memory()
        ref.accum_reset()
        torch.cuda.synchronize()
        for batch in range(2):
            expect_rows = []
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    nfr = 1 + (r + batch) % 2                      # ragged: ranks project different numbers of frames
                    for j in range(nfr):
                        pose = _frame_pose(10 * batch + 3 * r + j)[None]
                        n, h = ctxs[r].project_device(heat, K, pose, 0.5, "object", True, out=outs[r], sync=True)
                        ref.project_device(heat, K, pose, 0.5, "object", True, sync=True)
                    k = combs[r].acquire()
                    rec, cnt = combs[r].records(k)
                    ctxs[r].pack_records_device(outs[r]["t_hit"], outs[r]["face"], pixel=outs[r]["pixel"], n=n, out=rec,
                                                count_async=cnt, sync=False)
                    expect_rows.append(ctxs[r].pack_records_device(outs[r]["t_hit"], outs[r]["face"], pixel=outs[r]["pixel"], n=n).clone())
                    assert expect_rows[-1].shape[0] == h > 100
            for r in range(world):                                 # queued back to back: the kernels meet on the device
                with torch.cuda.stream(streams[r]):
                    combs[r].submit(gather_root=0, gathered=gathered, count_async=total_rows)
            torch.cuda.synchronize()
            for cb in combs:
                cb.check()
            hist, fmax, vmax = ref.accum_get()
            for cb in combs:
                a, b, c = cb.result()
                assert np.array_equal(a.cpu().numpy(), hist) and np.array_equal(b.cpu().numpy(), fmax) and np.array_equal(c.cpu().numpy(), vmax)
            want = torch.cat(expect_rows, dim=0)
            assert int(total_rows[0]) == want.shape[0]
            assert torch.equal(gathered[:want.shape[0]], want)
            for c in ctxs:                                         # the live blocks were zeroed by the snapshot
                assert not c.accum_get()[0].any()
    finally:
        for c in ctxs + [ref]:
            c.close()


@pytest.mark.parametrize("case", ["dense_tiled", "sparse"])
def test_sharded_frame_lands_in_every_window(built_lib, case, monkeypatch):
    """dp_peer_results: each rank traces its block of the frame and stores t_hit / face / point into the result slot of
    EVERY rank from the traversal's epilogue; after the frame barrier every window holds the unsharded frame bit for bit.
    Two frames (both slots), packets and the eight-lane sparse path, both node sets."""
    import torch
    from defectproj import Context, _lib
    from defectproj.projector import PeerCombiner
    V, F, H, W, K = _scene()
    if case == "dense_tiled":
        H, W = 256, 1024                                   # > 131072 rays: packets, walked in 8x4 tiles
        K = synth.K_matrix(126.0 * W / 512, 126.0 * W / 512, W / 2, H / 2)
        heat = torch.ones((1, H, W), device="cuda")
    else:
        heat = torch.from_numpy(synth.blob_heatmap((H, W), seed=5)).cuda()[None]
    n_px = H * W
    world = 3
    ctxs = [Context(0).set_mesh(V, F).build_bvh() for _ in range(world)]
    ref = Context(0).set_mesh(V, F).build_bvh()
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        combs = [PeerCombiner(c, r, world, result_rays=n_px, local_contexts=ctxs) for r, c in enumerate(ctxs)]
        for cb in combs:
            cb.open_local()
        for r, c in enumerate(ctxs):
            c.set_ray_shard(r, world)
        for fat in ("1", "0"):
            monkeypatch.setenv("DP_FAT", fat)
            for frame in range(2):
                pose = _frame_pose(frame)[None]
                whole = dict(t_hit=torch.empty(n_px, device="cuda"), face=torch.empty(n_px, dtype=torch.int32, device="cuda"),
                             point=torch.empty((n_px, 3), device="cuda"))
                ref.accum_reset()
                n, h = ref.project_device(heat, K, pose, 0.5, "object", True, out=whole, sync=True)
                slot = frame & 1
                for r, c in enumerate(ctxs):
                    c.accum_reset()
                    c.peer_results(slot, True)
                    with torch.cuda.stream(streams[r]):
                        c.project_device(heat, K, pose, 0.5, "object", True, sync=False)
                torch.cuda.synchronize()
                hs = 0
                for r, c in enumerate(ctxs):
                    assert c.peer_status() == 0
                    t = c.peer_tensor(_lib.DP_PEER_T_HIT, slot)[:n]
                    f = c.peer_tensor(_lib.DP_PEER_FACE, slot)[:n]
                    p = c.peer_tensor(_lib.DP_PEER_POINT, slot)[:n]
                    assert torch.equal(f, whole["face"][:n]), (fat, frame, r)
                    assert torch.equal(t.view(torch.int32), whole["t_hit"][:n].view(torch.int32))
                    assert torch.equal(p.view(torch.int32), whole["point"][:n].view(torch.int32))
                    hs = hs + c.accum_get()[0]
                assert np.array_equal(hs, ref.accum_get()[0]) and h > 100
        # argument checks of the mode
        with pytest.raises(ValueError):
            ctxs[0].project_device(heat, K, _frame_pose(0)[None], 0.5, "object", True, out=dict(t_hit=torch.empty(n_px, device="cuda")))
    finally:
        for c in ctxs + [ref]:
            c.close()


def test_missing_rank_times_out_instead_of_hanging(built_lib, monkeypatch):
    """A rank that never reaches the combine: the waiting kernel gives up after its 4 s budget, leaves the totals alone and
    sets the window's error word (PeerCombiner.check raises); the GPU is not left spinning."""
    import time
    import torch
    from defectproj import Context
    from defectproj.projector import PeerCombiner
    monkeypatch.setenv("DP_PEER_BLOCKS", "4")
    V, F, H, W, K = _scene()
    ctxs = [Context(0).set_mesh(V, F).build_bvh() for _ in range(2)]
    try:
        combs = [PeerCombiner(c, r, 2, local_contexts=ctxs) for r, c in enumerate(ctxs)]
        for cb in combs:
            cb.open_local()
        heat = torch.rand((1, H, W), device="cuda")
        ctxs[0].project_device(heat, K, _frame_pose(0)[None], 0.5, "object", True, sync=True)
        t0 = time.perf_counter()
        combs[0].submit()                                 # rank 1 never submits
        torch.cuda.synchronize()
        waited = time.perf_counter() - t0
        assert 3.0 < waited < 8.0
        assert ctxs[0].peer_status() == 1 and ctxs[1].peer_status() == 0
        with pytest.raises(RuntimeError):
            combs[0].check()
        assert not combs[0].result()[0].any()             # nothing was folded
    finally:
        for c in ctxs:
            c.close()


def test_peer_ranks_under_torchrun(built_lib):
    """The same over CUDA IPC between real ranks (needs >= 2 GPUs; the 1-GPU box skips)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the multi-process exchange is run by tests/tools/peer_ranks.py on a multi-GPU box")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "tools", "peer_ranks.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_RANKS_OK" in r.stdout
