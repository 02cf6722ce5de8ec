"""Multi-GPU semantics checked on ONE GPU (the N-rank NCCL path itself is exercised by bench.py --gpus N and by
tests/test_dist_cpu.py with gloo): single-frame ray sharding, the hit records a rank contributes to the gather,
the accumulator block and its per-batch combination."""
import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu


def _outs(n, point=False):
    import torch
    d = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), t_hit=torch.full((n,), -3.0, device="cuda"),
             face=torch.full((n,), -7, dtype=torch.int32, device="cuda"), intensity=torch.empty(n, device="cuda"))
    if point:
        d["point"] = torch.full((n, 3), -5.0, device="cuda")
    return d


@pytest.mark.parametrize("case", ["dense_tiled", "dense_odd", "sparse", "batch"])
def test_ray_shards_of_one_frame_concatenate_to_the_whole(ctx, case, monkeypatch):
    """dp_set_ray_shard: rank r traces block r of the compacted ray list; the slot ranges partition [0, n), nothing
    outside a rank's range is written, and slices / integer histograms / maxima combine to the unsharded result bit
    for bit -- for the tiled dense walk, a dense frame that cannot be tiled, a sparse frame (eight lanes per ray)
    and a batch of frames; with the uncompressed and the compressed node set."""
    import torch
    from defectproj import Context
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0, scale=6.0)
    pose = synth.fill_frame_pose()
    ctx.set_mesh(V, F).build_bvh()
    if case == "dense_tiled":
        H, W, B = 512, 512, 1                  # > 131072 rays: packets, walked in 8x4 tiles
        heat = torch.ones((B, H, W), device="cuda")
    elif case == "dense_odd":
        H, W, B = 510, 601, 1                  # cannot be tiled: packets in row-major order
        heat = torch.ones((B, H, W), device="cuda")
    elif case == "sparse":
        H, W, B = 360, 640, 1
        heat = torch.from_numpy(synth.blob_heatmap((H, W), seed=5)).cuda()[None]
    else:
        H, W, B = 128, 256, 3
        heat = torch.rand((B, H, W), device="cuda")
    K = synth.K_matrix(126.0 * W / 512, 126.0 * W / 512, W / 2, H / 2)
    poses = np.stack([pose] * B)
    n_px = B * H * W
    for fat in ("1", "0"):
        monkeypatch.setenv("DP_FAT", fat)
        ctx.set_ray_shard(0, 1)
        ctx.accum_reset()
        whole = _outs(n_px, point=True)
        n, h = ctx.project_device(heat, K, poses, 0.5, "object", True, out=whole, sync=True)
        hw, fw, vw = ctx.accum_get()
        assert h > 1000
        for world in (2, 3, 8):
            hs, fs, vs, hits, covered = [], [], [], 0, 0
            merged = _outs(n_px, point=True)
            for r in range(world):
                ctx.set_ray_shard(r, world)
                ctx.accum_reset()
                part = _outs(n_px, point=True)
                n_r, h_r = ctx.project_device(heat, K, poses, 0.5, "object", True, out=part, sync=True)
                lo, hi = Context.shard_slots(r, world, n_r, H, W, nframes=B)
                assert n_r == n and lo == covered and hi >= lo
                covered = hi
                hits += h_r
                # untouched outside the shard
                assert (part["face"][:lo] == -7).all() and (part["face"][hi:n] == -7).all()
                assert (part["t_hit"][:lo] == -3.0).all() and (part["t_hit"][hi:n] == -3.0).all()
                for k in ("t_hit", "face", "point"):
                    merged[k][lo:hi] = part[k][lo:hi]
                assert torch.equal(part["pixel"][:n], whole["pixel"][:n])     # the selection is replicated
                a, b, c = ctx.accum_get()
                hs.append(a); fs.append(b); vs.append(c)
            assert covered == n and hits == h
            assert torch.equal(merged["face"][:n], whole["face"][:n])
            assert torch.equal(merged["t_hit"][:n].view(torch.int32), whole["t_hit"][:n].view(torch.int32))
            assert torch.equal(merged["point"][:n].view(torch.int32), whole["point"][:n].view(torch.int32))
            assert np.array_equal(np.sum(hs, axis=0, dtype=np.int32), hw)
            assert np.array_equal(np.max(fs, axis=0), fw) and np.array_equal(np.max(vs, axis=0), vw)
    ctx.set_ray_shard(0, 1)
    with pytest.raises(ValueError):
        ctx.set_ray_shard(2, 2)


def test_hit_records_are_the_hits_in_ray_order(ctx):
    """dp_pack_records: rows (pixel, t bits, face[, point bits]) of exactly the rays that hit
    (/root/reference/src/defect_projection.py:259-264), in ray order; sub-ranges; the asynchronous count."""
    import torch
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    ctx.set_mesh(V, F).build_bvh()
    heat = torch.from_numpy(synth.blob_heatmap((H, W), seed=4)).cuda()[None]
    out = _outs(H * W, point=True)
    n, h = ctx.project_device(heat, K, synth.fixed_pose()[None], 0.3, "object", False, out=out, sync=True)
    assert 0 < h < n
    face = out["face"][:n].cpu().numpy()
    keep = face >= 0
    rec = ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n).cpu().numpy()
    assert rec.shape == (h, 3)
    assert np.array_equal(rec[:, 0], out["pixel"][:n].cpu().numpy()[keep])
    assert np.array_equal(rec[:, 1], out["t_hit"][:n].view(torch.int32).cpu().numpy()[keep])
    assert np.array_equal(rec[:, 2], face[keep])
    rec6 = ctx.pack_records_device(out["t_hit"], out["face"], point=out["point"], n=n).cpu().numpy()
    assert rec6.shape == (h, 6)
    assert np.array_equal(rec6[:, 0], np.nonzero(keep)[0].astype(np.int32))           # no pixel list: the ray index
    assert np.array_equal(rec6[:, 3:], out["point"][:n].view(torch.int32).cpu().numpy()[keep])
    # a sub-range (a rank's shard), count delivered asynchronously to pinned memory
    lo, hi = n // 3, n // 3 + 5000
    cnt = torch.zeros(1, dtype=torch.int64).pin_memory()
    buf = torch.empty((hi - lo, 3), dtype=torch.int32, device="cuda")
    ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], first=lo, n=hi - lo, out=buf, count_async=cnt,
                            sync=False)
    torch.cuda.synchronize()
    m = int(cnt[0])
    assert m == keep[lo:hi].sum()
    assert np.array_equal(buf[:m, 2].cpu().numpy(), face[lo:hi][keep[lo:hi]])
    # nothing hits / nothing to pack
    none = ctx.pack_records_device(out["t_hit"], torch.full_like(out["face"], -1), n=n)
    assert none.shape[0] == 0
    assert ctx.pack_records_device(out["t_hit"], out["face"], n=0).shape[0] == 0


def test_batch_combiner_counts_every_hit_once(ctx):
    """BatchCombiner on one rank: two batches accumulated through snapshots == one accumulation of both; the live
    accumulators are left zeroed (local), never reduced in place; the accumulator block is one allocation."""
    import torch
    from defectproj.projector import BatchCombiner
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["small"], seed=2)
    K, H, W = synth.K_matrix(150.0, 150.0, 80.0, 60.0), 120, 160
    poses = synth.fibonacci_poses(6, radius=400.0)
    heats = torch.from_numpy(np.stack([synth.blob_heatmap((H, W), seed=i) for i in range(6)])).cuda()
    ctx.set_mesh(V, F).build_bvh()
    base, h_off, f_off, v_off, nbytes = ctx.accum_layout()
    hp, fp, vp = ctx.accum_device_ptrs()
    assert (hp, fp, vp) == (base + h_off, base + f_off, base + v_off) and f_off % 256 == 0 and v_off % 256 == 0
    assert nbytes >= v_off + 4 * len(V)
    ctx.accum_reset()
    ctx.project_device(heats, K, poses, 0.4, "object", True, sync=True)
    h0, f0, v0 = ctx.accum_get()
    assert h0.sum() > 500
    comb = BatchCombiner(ctx)
    ctx.accum_reset()
    for lo, hi in ((0, 2), (2, 6)):
        ctx.project_device(heats[lo:hi], K, poses[lo:hi], 0.4, "object", True, sync=False)
        comb.submit()
    h, f, v = comb.result()
    torch.cuda.synchronize()
    assert np.array_equal(h.cpu().numpy(), h0) and np.array_equal(f.cpu().numpy(), f0) and np.array_equal(v.cpu().numpy(), v0)
    assert ctx.accum_get()[0].sum() == 0
    comb.reset_totals()
    assert int(comb.result()[0].sum()) == 0


def test_sharded_frame_api_on_one_rank_and_the_early_ray_count(built_lib):
    """Projector.project_frame_sharded without a process group (world 1): the ray count is read from pinned memory as soon
    as the compaction kernel has stored it, nothing waits for the traversal, results equal the plain projection."""
    import torch
    from defectproj import Projector
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0, scale=6.0)
    pose = synth.fill_frame_pose()
    H, W = 384, 512
    K = synth.K_matrix(126.0, 126.0, W / 2, H / 2)
    proj = Projector(V, F, device=0)
    try:
        for heat in (torch.ones((H, W), device="cuda"), torch.from_numpy(synth.blob_heatmap((H, W), seed=3)).cuda(),
                     torch.zeros((H, W), device="cuda")):
            out = _outs(H * W)
            n, h, (lo, hi) = proj.project_frame_sharded(heat, K, pose, 0.5, out={"t_hit": out["t_hit"], "face": out["face"]})
            ref = _outs(H * W)
            proj.ctx.accum_reset()
            n0, h0 = proj.ctx.project_device(heat[None], K, pose[None], 0.5, "object", True, out=ref, sync=True)
            assert (n, int(h), lo, hi) == (n0, h0, 0, n0)
            assert torch.equal(out["face"][:n], ref["face"][:n])
            assert torch.equal(out["t_hit"][:n].view(torch.int32), ref["t_hit"][:n].view(torch.int32))
            assert int(proj.combined()[0].sum()) == h0
    finally:
        proj.ctx.close()
