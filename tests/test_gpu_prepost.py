"""GPU parity for the steps either side of the back-projection:
heatmap preparation (DataReader.get_heatmap, datareader.py:639-675), the viewer payload
(create_intersection_pcd :268-294 + hit selection :259-264 + PointCloud.transform run.py:118) and the
multi-frame accumulation loop of run.py:175-206.  Bars: bit-exact float64 / float32 against the golden
fixtures produced by the reference's own code (real cv2) and against the oracle on seeded inputs."""
import os

import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def depth_golden():
    return np.load(os.path.join(HERE, "golden", "depth_path.npz"))


# ------------------------------------------------------------------------------------------ 8f #3
def test_prepare_heatmap_equals_reference_get_heatmap(ctx, depth_golden):
    g = depth_golden
    for tag in "abcd":
        cH, cW, ds = (int(v) for v in g[f"h_cfg_{tag}"])
        H, W = int(cH / ds), int(cW / ds)
        out = ctx.prepare_heatmap(g[f"h_data_{tag}"], H, W)
        assert out.dtype == np.float64 and out.shape == (H, W)
        assert np.array_equal(out, g[f"h_full_{tag}"]), tag                 # bit-exact with cv2's output
        out32 = ctx.prepare_heatmap(g[f"h_data_{tag}"], H, W, np.float32)
        assert np.array_equal(out32, g[f"h_full_{tag}"].astype(np.float32))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(50, 70, 120, 200), (224, 224, 720, 1280), (33, 21, 64, 48), (300, 300, 77, 91),
                                   (7, 7, 1, 5), (1, 1, 16, 16), (224, 224, 1024, 1024), (64, 64, 2160, 3840)])
def test_prepare_heatmap_equals_oracle(ctx, orc, dt, shape):
    sh, sw, H, W = shape
    rng = np.random.default_rng(sh * 7 + W)
    data = (rng.random((sh, sw)) * 5 - 2).astype(dt)
    if sh * sw == 1:
        with np.errstate(invalid="ignore"):
            assert np.array_equal(ctx.prepare_heatmap(data, H, W), orc.prepare_heatmap(data, H, W), equal_nan=True)
        return
    assert np.array_equal(ctx.prepare_heatmap(data, H, W), orc.prepare_heatmap(data, H, W))


def test_prepare_heatmap_edge_cases(ctx, orc):
    # constant map: 0/0 everywhere inside the window, zeros outside -- what numpy gives the reference
    d = np.full((8, 8), 3.0)
    out = ctx.prepare_heatmap(d, 12, 20)
    assert np.isnan(out[:, 4:16]).all() and (out[:, :4] == 0).all() and (out[:, 16:] == 0).all()
    # NaN in the data poisons min/max like np.min / np.max
    d = np.arange(64, dtype=np.float64).reshape(8, 8)
    d[3, 3] = np.nan
    assert np.isnan(ctx.prepare_heatmap(d, 8, 8)).all()
    # values stay within [0, 1] and the extremes are reached when the size is kept
    d = np.random.default_rng(0).random((32, 32))
    out = ctx.prepare_heatmap(d, 32, 32)
    assert out.min() == 0.0 and out.max() == 1.0
    assert np.array_equal(out, (d - d.min()) / (d - d.min()).max())        # same size: the resize is the identity
    with pytest.raises(ValueError):
        ctx.prepare_heatmap(np.zeros((0, 4)), 8, 8)
    # integer maps are promoted like numpy does (int - int -> int, / -> float64)
    di = np.random.default_rng(1).integers(0, 255, (20, 20))
    assert np.array_equal(ctx.prepare_heatmap(di, 30, 40), orc.prepare_heatmap(di.astype(np.float64), 30, 40))


def test_prepared_heatmap_feeds_the_projection(ctx, orc):
    """get_heatmap -> ray_tracing chain with the heatmap never leaving the device."""
    torch = pytest.importorskip("torch")
    V, F = synth.param_mesh(40, 25, seed=4)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    ctx.set_mesh(V, F)
    ctx.build_bvh()
    data = np.random.default_rng(5).random((224, 224)).astype(np.float32) * synth.gaussian_heatmap((224, 224), sigma=60.0).astype(np.float32)
    d_data = torch.from_numpy(data).cuda()
    d_heat = ctx.prepare_heatmap(d_data, H, W, np.float32)
    assert d_heat.is_cuda and tuple(d_heat.shape) == (H, W)
    ref_heat = orc.prepare_heatmap(data, H, W).astype(np.float32)
    assert np.array_equal(d_heat.cpu().numpy(), ref_heat)
    res = ctx.project(ref_heat, K, pose[None], 0.75, frame="object", accumulate=False, want=("pixel", "face"))
    pix, _, _ = ctx.compact(d_heat.cpu().numpy(), 0.75)
    assert np.array_equal(res["pixel"], pix) and res["n"] == (ref_heat > np.float32(0.75)).sum()


def test_heatmap_reader_mirrors_get_heatmap(tmp_path, depth_golden):
    from defectproj.datareader import HeatmapReader
    g = depth_golden
    os.makedirs(tmp_path / "heatmap")
    np.save(tmp_path / "heatmap" / "0002.npy", g["h_data_c"])
    cH, cW, ds = (int(v) for v in g["h_cfg_c"])
    color = np.zeros((cH, cW, 3), np.uint8)
    full, color_original, vis, again = HeatmapReader(str(tmp_path), cH, cW, ds).get_heatmap(color)
    assert np.array_equal(full, g["h_full_c"])
    o = min(full.shape)
    assert vis.shape == (o, o) and (color_original is None or color_original.shape[:2] == (o, o))
    assert again is color_original


def test_device_resident_heatmap_gives_the_host_array_result(tmp_path, depth_golden, golden):
    """HeatmapReader.get_heatmap(device=True) keeps heatmap_full on the GPU (bit-identical to the host version, which is
    the reference's own cv2 output); ray_tracing() takes that CUDA tensor and returns exactly what it returns for the
    host array: hit cloud, colours, face ids, pixels, t_hit, posed mesh, accumulators, miss-all LineSet."""
    import torch
    from defectproj import defect_projection as dpj
    from defectproj.datareader import HeatmapReader
    g = depth_golden
    os.makedirs(tmp_path / "heatmap")
    for tag in ("c", "d"):                                          # float64 and float32 raw maps
        np.save(tmp_path / "heatmap" / "0002.npy", g[f"h_data_{tag}"])
        cH, cW, ds = (int(v) for v in g[f"h_cfg_{tag}"])
        rd = HeatmapReader(str(tmp_path), cH, cW, ds)
        full_h, _, vis_h, _ = rd.get_heatmap()
        full_d, _, vis_d, _ = rd.get_heatmap(device=True)
        assert full_d.is_cuda and np.array_equal(full_d.cpu().numpy(), full_h) and np.array_equal(full_h, g[f"h_full_{tag}"])
        assert np.array_equal(vis_d.cpu().numpy(), vis_h) and vis_d.cpu().numpy().dtype == vis_h.dtype
    K = golden["g5_K"]
    synth.write_scene_dir(str(tmp_path), K, (96, 128), color_to_depth=golden["g5_color_to_depth"])
    mesh = dpj.TriangleMesh(golden["g5_V_depthcam"], golden["g5_F"])
    for heat in (golden["g2_heat"], golden["g2_heat"].astype(np.float32)):
        for thr in (0.5, 0.75):
            a, ma = dpj.ray_tracing(str(tmp_path), mesh, heat, K, heatmap_threshold=thr)
            la = dict(dpj.last_result())
            fa = dpj.face_intensities()
            b, mb = dpj.ray_tracing(str(tmp_path), mesh, torch.from_numpy(heat).cuda(), K, heatmap_threshold=thr)
            lb = dict(dpj.last_result())
            fb = dpj.face_intensities()
            assert len(a.points) == len(b.points) > 0
            for k in ("points", "colors", "face_ids", "pixels", "t_hit"):
                assert np.array_equal(getattr(a, k), getattr(b, k)), k
            assert np.array_equal(ma.vertices, mb.vertices)
            for k in ("pixel", "intensity", "t_hit", "face", "n_rays", "n_hits"):
                assert np.array_equal(la[k], lb[k]), k
            assert la["intensity"].dtype == lb["intensity"].dtype
            for x, y in zip(fa, fb):
                assert np.array_equal(x, y)
    far = dpj.TriangleMesh(golden["g5_V_model"].astype(np.float64) + np.array([5000.0, 0, 0]), golden["g5_F"])
    ls_h, _ = dpj.ray_tracing(str(tmp_path), far, golden["g2_heat"], K, heatmap_threshold=0.9)
    ls_d, _ = dpj.ray_tracing(str(tmp_path), far, torch.from_numpy(golden["g2_heat"]).cuda(), K, heatmap_threshold=0.9)
    assert isinstance(ls_d, dpj.LineSet) and np.array_equal(ls_d.points, ls_h.points) and np.array_equal(ls_d.lines, ls_h.lines)
    ls_e, _ = dpj.ray_tracing(str(tmp_path), mesh, torch.zeros((96, 128), device="cuda"), K)
    assert len(ls_e.lines) == 0


# ------------------------------------------------------------------------------------------ 8f #2
def test_colours_equal_reference_create_intersection_pcd(ctx, golden):
    from defectproj import defect_projection as dpj
    pcd = dpj.create_intersection_pcd(np.zeros((33, 3)), golden["g7_ramp"])
    assert np.array_equal(pcd.colors, golden["g7_colors"])                  # made by the reference's function
    const = dpj.create_intersection_pcd(np.zeros((4, 3)), np.full(4, 0.7))
    assert np.array_equal(const.colors, np.zeros((4, 3)))                   # matplotlib's 'bad' colour for 0/0
    assert len(dpj.create_intersection_pcd(np.zeros((0, 3)), np.zeros(0)).colors) == 0
    assert np.array_equal(ctx.jet_lut()[[0, 255]], [[0, 0, 0.5], [0.5, 0, 0]])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 100_003, 1_048_576])
def test_pack_hits_equals_oracle(ctx, orc, dt, n):
    rng = np.random.default_rng(n)
    I = rng.random(n).astype(dt)
    face = rng.integers(-1, 50, n).astype(np.int32)
    face[rng.random(n) < 0.4] = -1
    pix = rng.permutation(n).astype(np.uint32)
    p64 = rng.normal(size=(n, 3)) * 100
    T = np.eye(4)
    T[:3, :3] = synth.rot_z(33.0) @ synth.rot_x(-12.0)
    T[:3, 3] = [-32.0, -2.0, 4.0]
    for Tm in (None, T):
        got = ctx.pack_hits(I, face, pix, p64, T=Tm)
        ref = orc.pack_hits(I, face, p64, Tm)
        assert got["m"] == len(ref["index"])
        assert np.array_equal(got["pixel"], pix[ref["index"]])              # ray order is kept
        assert np.array_equal(got["face"], face[ref["index"]])
        assert np.array_equal(got["intensity"], ref["intensity"])
        assert np.array_equal(got["colors"], ref["colors"])                 # bit-exact float64 colours
        assert np.array_equal(got["points"], ref["points"])                 # same stated order of operations
    allsel = ctx.pack_hits(I)                                               # face=None: every ray
    assert allsel["m"] == n and np.array_equal(allsel["colors"], orc.pack_hits(I)["colors"])
    none = ctx.pack_hits(I, np.full(n, -1, np.int32), pix, p64)
    assert none["m"] == 0 and none["points"].shape == (0, 3)


def test_pack_hits_device_resident_chain(ctx, orc):
    """project_device -> pack_hits_device: the viewer payload is built from the per-ray device buffers; only the
    counts cross PCIe.  Equals the host-buffer path and the oracle."""
    torch = pytest.importorskip("torch")
    V, F = synth.param_mesh(40, 25, seed=4)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    ctx.set_mesh(V, F).build_bvh()
    heat_np = synth.gaussian_heatmap((H, W), dtype=np.float32)
    heat = torch.from_numpy(heat_np)[None].cuda()
    n = H * W
    o = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), intensity=torch.empty(n, device="cuda"),
             face=torch.empty(n, dtype=torch.int32, device="cuda"), point64=torch.empty((n, 3), dtype=torch.float64, device="cuda"))
    nr, nh = ctx.project_device(heat, K, pose[None], 0.5, "object", False, out=o, sync=True)
    T = np.eye(4)
    T[:3, :3] = synth.rot_y(1.5)
    T[:3, 3] = [-32.0, -2.0, 4.0]
    dev = ctx.pack_hits_device(o["intensity"][:nr], o["face"][:nr], o["pixel"][:nr], o["point64"][:nr], T=T)
    assert dev["m"] == nh and all(v.is_cuda for k, v in dev.items() if k != "m")
    host = ctx.pack_hits(o["intensity"][:nr].cpu().numpy(), o["face"][:nr].cpu().numpy(),
                         o["pixel"][:nr].cpu().numpy().view(np.uint32), o["point64"][:nr].cpu().numpy(), T=T)
    ref = orc.pack_hits(o["intensity"][:nr].cpu().numpy(), o["face"][:nr].cpu().numpy(), o["point64"][:nr].cpu().numpy(), T)
    for k in ("points", "colors", "face", "intensity"):
        assert np.array_equal(dev[k].cpu().numpy(), host[k]), k
    assert np.array_equal(dev["pixel"].cpu().numpy().view(np.uint32), host["pixel"])
    assert np.array_equal(host["colors"], ref["colors"]) and np.array_equal(host["points"], ref["points"])
    with pytest.raises(ValueError):
        ctx.pack_hits_device(o["intensity"][:nr], o["face"][:nr - 1])


def test_pack_hits_capacity_and_arguments(ctx, built_lib):
    import ctypes as C
    I = np.ones(16, np.float32)
    col = np.empty((4, 3), np.float64)
    m = C.c_int64(0)
    vp = C.c_void_p
    rc = built_lib.dp_pack_hits(ctx._h, I.ctypes.data_as(vp), 0, None, None, None, 16, None, None, col.ctypes.data_as(vp),
                                None, None, None, 4, C.byref(m), 0, None)
    assert rc == -3 and m.value == 16                                       # DP_E_NOMEM, count still reported
    rc = built_lib.dp_pack_hits(ctx._h, None, 0, None, None, None, 16, None, None, None, None, None, None, 0, C.byref(m), 0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        ctx.pack_hits(np.ones(4), face=np.zeros(5, np.int32))


def test_transform_points_equals_oracle_and_open3d_semantics(ctx, orc):
    rng = np.random.default_rng(9)
    p = rng.normal(size=(50_001, 3)) * 300
    T = np.eye(4)
    T[:3, :3] = synth.rot_y(41.0) @ synth.rot_z(-7.0)
    T[:3, 3] = [5.0, 6.0, -700.0]
    q = ctx.transform_points(p, T)
    assert np.array_equal(q, orc.transform_points(p, T))
    assert np.allclose(q, p @ T[:3, :3].T + T[:3, 3], rtol=0, atol=1e-9)
    Tp = T.copy()
    Tp[3] = [0, 0, 0, 2.0]                                                   # homogeneous divide like Open3D
    assert np.array_equal(ctx.transform_points(p, Tp), orc.transform_points(p, Tp))
    from defectproj import defect_projection as dpj
    pcd = dpj.PointCloud(p)
    assert pcd.transform(T) is pcd and np.array_equal(pcd.points, q)


def test_ray_tracing_payload_for_the_viewer(ctx, tmp_path, golden):
    """ray_tracing -> transform -> update_dash_data, the sequence of run.py:113-119 / :205."""
    from defectproj import defect_projection as dpj, web_vis
    g = golden
    K = g["g5_K"]
    synth.write_scene_dir(str(tmp_path), K, (96, 128), color_to_depth=g["g5_color_to_depth"])
    mesh = dpj.TriangleMesh(g["g5_V_depthcam"], g["g5_F"])
    pcd, posed = dpj.ray_tracing(str(tmp_path), mesh, g["g2_heat"], K, heatmap_threshold=0.5)
    assert np.allclose(pcd.points, g["g5_points_050"], rtol=0, atol=1e-9)
    assert np.array_equal(pcd.colors, g["g5_colors_050"])
    before = pcd.points.copy()
    pcd.transform(g["g5_color_to_depth"])

    class Q:
        def put(self, item):
            self.item = item
    q = Q()
    payload = web_vis.update_dash_data([pcd], posed, queue=q)
    assert q.item is payload and set(payload) >= {"pcds", "vertices", "faces", "face_hits", "face_intensity"}
    assert np.array_equal(payload["pcds"][0]["points"], pcd.points) and not np.array_equal(before, pcd.points)
    assert payload["face_hits"].sum() == len(pcd.points)
    assert payload["faces"].shape == (len(g["g5_F"]), 3) and payload["vertices"].shape == (len(g["g5_V_model"]), 3)
    hit_faces = np.unique(pcd.face_ids)
    assert (payload["face_intensity"][hit_faces] > 0.5).all()
    kw = web_vis.mesh3d_kwargs(payload)
    assert kw["intensitymode"] == "cell" and len(kw["intensity"]) == len(g["g5_F"])


def test_defect_tracker_follows_the_run_loop(ctx, orc):
    """run.py:100-120 + :175-206 restated with numpy/oracle pieces, against DefectTracker."""
    from defectproj.tracking import DefectTracker
    V, F = synth.param_mesh(40, 25, seed=4)
    V = V.astype(np.float64)
    K, H, W = synth.camera_720p()
    c2d = np.eye(4)
    c2d[:3, :3] = synth.rot_y(1.5)
    c2d[:3, 3] = [-32.0, -2.0, 4.0]
    trk = DefectTracker((V, F), K, c2d, heatmap_threshold=0.75)
    poses = [synth.fixed_pose(), synth.fixed_pose() @ synth._pose(synth.rot_z(10.0), [3.0, -2.0, 5.0]),
             synth.fixed_pose() @ synth._pose(synth.rot_x(-8.0), [0.0, 4.0, -6.0])]
    heats = [synth.gaussian_heatmap((H, W), sigma=40.0 + 10 * i) for i in range(3)]
    ref_clouds, prev = [], None
    hist_total = np.zeros(len(F), np.int64)
    for heat, pose in zip(heats, poses):
        cur = np.linalg.inv(pose)                      # ICP result: camera(depth) -> model
        pcd = trk.add_detection(heat, cur)
        # reference loop, on the CPU
        T = np.linalg.inv(c2d) @ np.linalg.inv(cur)
        Vc = orc.pose_vertices(V, T)
        bvh = orc.Bvh(Vc, F)
        xs, ys, I = orc.heatmap_to_points(heat, 0.75)
        rays = orc.compute_rays(xs, ys, K)
        t, f = bvh.cast_f32(orc.rays6_camera(rays))
        hit = f >= 0
        pts = rays[hit] * t[hit, None].astype(np.float64)
        if prev is not None:
            rel = np.linalg.inv(cur) @ prev
            ref_clouds = [orc.transform_points(c, rel) for c in ref_clouds]
        ref_clouds.append(orc.transform_points(pts, c2d))
        prev = cur
        hist_total += np.bincount(f[hit], minlength=len(F))
        assert np.array_equal(pcd.face_ids, f[hit])
    assert len(trk.intersection_pcds) == 3
    for got, ref in zip(trk.intersection_pcds, ref_clouds):
        assert np.array_equal(got.points, ref)
    assert np.array_equal(trk.hist, hist_total)
    p = trk.payload()
    assert len(p["pcds"]) == 3 and p["face_hits"].sum() == sum(len(c) for c in ref_clouds)
    # the payload's mesh is the model posed by inv(current) only -- the depth camera's frame, the frame of the clouds
    # (run.py:179-181, :205) -- so the newest cloud's points lie ON the payload mesh: each in the plane and inside the
    # triangle of its face
    Vd = V @ np.linalg.inv(prev)[:3, :3].T + np.linalg.inv(prev)[:3, 3]
    assert np.allclose(p["vertices"], Vd, rtol=0, atol=1e-9)
    last = trk.intersection_pcds[-1]
    tri = p["vertices"][F[last.face_ids]]                                   # [m, 3, 3]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dist = np.abs(np.einsum("ij,ij->i", last.points - tri[:, 0], nrm))
    assert dist.max() < 1e-3                                                # mm; float32 t_hit at ~600 mm
