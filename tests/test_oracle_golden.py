"""CPU: the oracle (oracle/oracle.c) against the fixtures made from the reference's own Python
(tests/golden/make_golden.py) and against independent restatements."""
import numpy as np
import pytest

from defectproj import synth


def test_heatmap_to_points_known_answer(orc, golden):
    # hand-derivable example of SURVEY.md section 4
    xs, ys, I = orc.heatmap_to_points(golden["g1_heat"], 0.5)
    got = np.stack([xs, ys, I], axis=1)
    assert np.array_equal(got, golden["g1_points"])
    assert got.tolist() == [[3.0, 0.0, 0.9], [0.0, 1.0, 0.8], [1.0, 2.0, 0.51]]


@pytest.mark.parametrize("thr,key", [(0.5, "g2_points_050"), (0.75, "g2_points_075")])
def test_heatmap_to_points_reference_gaussian(orc, golden, thr, key):
    xs, ys, I = orc.heatmap_to_points(golden["g2_heat"], thr)
    ref = golden[key]
    assert np.array_equal(xs, ref[:, 0].astype(np.int64))
    assert np.array_equal(ys, ref[:, 1].astype(np.int64))
    assert np.array_equal(I, ref[:, 2])          # bit-exact float64


def test_production_heatmap_counts(golden):
    # the counts the survey measured with the reference's own generator (cv2 GaussianBlur)
    assert golden["g3_counts"].tolist() == [10885, 4501]


def test_compute_rays_bit_exact(orc, golden):
    ref = golden["g2_points_050"]
    rays = orc.compute_rays(ref[:, 0].astype(np.int64), ref[:, 1].astype(np.int64), golden["g4_K"])
    assert np.array_equal(rays, golden["g4_rays"])
    assert np.array_equal(ref[:, 2], golden["g4_intensities"])


def test_pose_vertices_matches_reference_transform(orc, golden):
    T = np.linalg.inv(golden["g5_color_to_depth"])
    got = orc.pose_vertices(golden["g5_V_depthcam"], T)
    assert np.array_equal(got, golden["g5_V_colorcam_050"].astype(np.float32))


@pytest.mark.parametrize("thr,tag", [(0.5, "050"), (0.75, "075")])
def test_whole_path_against_reference_run(orc, golden, thr, tag):
    """ray_tracing() of the reference (open3d shimmed) vs the oracle pipeline."""
    T = np.linalg.inv(golden["g5_color_to_depth"])
    Vc = orc.pose_vertices(golden["g5_V_depthcam"], T)
    xs, ys, I = orc.heatmap_to_points(golden["g2_heat"], thr)
    d = orc.compute_rays(xs, ys, golden["g5_K"])
    bvh = orc.Bvh(Vc, golden["g5_F"])
    t, f = bvh.cast_f32(orc.rays6_camera(d))
    hit = f >= 0
    pts = d[hit] * t[hit].astype(np.float64)[:, None]
    assert np.array_equal(pts, golden[f"g5_points_{tag}"])


def test_oracle_bvh_equals_brute_force(orc):
    V, F = synth.param_mesh(40, 25, seed=4)
    K, H, W = synth.camera_720p()
    ys, xs = np.mgrid[0:H:16, 0:W:16]
    xf = orc.frame_xform(K, synth.fixed_pose())
    rays6 = orc.rays_object_frame(xs.reshape(-1), ys.reshape(-1), xf)
    t0, f0 = orc.cast_brute_f32(V, F, rays6)
    t1, f1 = orc.Bvh(V, F).cast_f32(rays6)
    assert (f0 >= 0).sum() > 50
    assert np.array_equal(f0, f1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_float32_contract_agrees_with_float64_on_non_tie_rays(orc):
    V, F = synth.param_mesh(24, 16, seed=7)
    K, H, W = synth.camera_720p()
    ys, xs = np.mgrid[0:H:8, 0:W:8]
    xf = orc.frame_xform(K, synth.fixed_pose(z=420.0))
    rays6 = orc.rays_object_frame(xs.reshape(-1), ys.reshape(-1), xf)
    bvh = orc.Bvh(V, F)
    t32, f32 = bvh.cast_f32(rays6)
    t64, f64, tie = bvh.cast_f64(rays6)
    tn, fn = orc.brute_f64_numpy(V, F, rays6)
    clean = tie == 0
    assert clean.sum() > 0.9 * len(tie)
    assert np.array_equal(f64[clean], fn[clean])           # C float64 == numpy float64
    assert np.array_equal(f32[clean], f64[clean])          # float32 contract == float64 truth off ties
    hit = clean & (f64 >= 0)
    diag = np.linalg.norm(V.max(0) - V.min(0))
    d = rays6[:, 3:].astype(np.float64)
    err = np.linalg.norm(d[hit] * (t32[hit].astype(np.float64) - t64[hit])[:, None], axis=1)
    assert err.max() <= 1e-5 * diag


def test_single_triangle_known_answer(orc):
    v0, v1, v2 = [-1, -1, 5], [1, -1, 5], [0, 1, 5]
    hit, t = orc.tri_test_f32([0, 0, 0, 0, 0, 1], v0, v1, v2)
    assert hit and t == 5.0
    assert not orc.tri_test_f32([0, 0, 0, 0, 0, -1], v0, v1, v2)[0]     # t >= 0 only
    assert not orc.tri_test_f32([5, 5, 0, 0, 0, 1], v0, v1, v2)[0]


def test_accumulate_conservation(orc):
    rng = np.random.default_rng(0)
    F = rng.integers(0, 50, (30, 3)).astype(np.int32)
    face = rng.integers(-1, 30, 1000).astype(np.int32)
    I = rng.random(1000).astype(np.float32)
    hist, fmax, vmax = orc.accumulate(face, I, F, 50)
    assert hist.sum() == (face >= 0).sum()
    assert fmax.max() <= I.max() and vmax.max() == fmax.max()
    for f in range(30):
        sel = face == f
        assert hist[f] == sel.sum()
        assert fmax[f] == (I[sel].max() if sel.any() else 0.0)


# ------------------------------------------------------------------ depth-image projection path (8f #1)
@pytest.fixture(scope="module")
def depth_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_path.npz"))


def test_depth_path_oracle_vs_reference(orc, depth_golden):
    g = depth_golden
    for thr, key in ((0.1, "d_point3d_010"), (0.5, "d_point3d_050")):
        assert np.array_equal(orc.heatmap_to_point3d(g["d_heat"], g["d_depth"], g["d_K"], thr), g[key])
    assert np.array_equal(orc.heatmap_to_point3d(g["d_heat"], g["d_depth"][:50, :70], g["d_K"], 0.3), g["d_point3d_small"])
    # float32 map: float32 division, the rounded quotient in column 3 (fixture made by the reference's own function)
    for thr, key in ((0.1, "d_point3d_f32_010"), (0.5, "d_point3d_f32_050")):
        got = orc.heatmap_to_point3d(g["d_heat"].astype(np.float32), g["d_depth"], g["d_K"], thr)
        assert np.array_equal(got, g[key])
    assert not np.array_equal(g["d_point3d_f32_050"][:, 3], g["d_point3d_050"][:, 3])      # the float32 quotients differ
    offs, ali, idx = orc.align_to_surface(g["d_proj_point3d"], g["d_target_points"], g["d_target_normals"], 0.5)
    assert np.array_equal(offs, g["d_proj_offset"]) and np.array_equal(ali, g["d_proj_aligned"])
    offs, ali, _ = orc.align_to_surface(g["d_point3d_050"], g["d_target_points"], g["d_target_normals"], 0.1)
    assert np.array_equal(offs, g["d_align_offset_01"]) and np.array_equal(ali, g["d_align_aligned_01"])


# ------------------------------------------------------------------ heatmap preparation (8f #3)
def test_prepare_heatmap_oracle_vs_reference_get_heatmap(orc, depth_golden):
    """h_full_* were produced by the reference's DataReader.get_heatmap run verbatim with the real cv2."""
    g = depth_golden
    for tag in "abcd":
        cH, cW, ds = (int(v) for v in g[f"h_cfg_{tag}"])
        out = orc.prepare_heatmap(g[f"h_data_{tag}"], int(cH / ds), int(cW / ds))
        assert out.dtype == np.float64 and np.array_equal(out, g[f"h_full_{tag}"]), tag
    assert g["h_data_d"].dtype == np.float32          # the CV_32F path is covered


def test_prepare_heatmap_oracle_vs_live_cv2(orc):
    """Where cv2 is importable the restatement is also checked live (the wheel's IPP path may differ per CPU)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64):
        for sh, sw, H, W in ((50, 70, 120, 200), (224, 224, 720, 1280), (33, 21, 64, 48), (300, 300, 77, 91), (7, 7, 1, 5)):
            data = (rng.random((sh, sw)) * 3 - 1).astype(dt)
            h = data - np.min(data)
            h = h / np.max(h)
            o = min(H, W)
            full = np.zeros((H, W))
            y0, x0 = (H - o) // 2, (W - o) // 2
            full[y0:y0 + o, x0:x0 + o] = cv2.resize(h, (o, o), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(orc.prepare_heatmap(data, H, W), full), (dt, sh, sw, H, W)


def test_transform_points_oracle_is_a_rigid_map(orc):
    from defectproj import synth
    rng = np.random.default_rng(2)
    p = rng.normal(size=(100, 3)) * 50
    T = np.eye(4)
    T[:3, :3] = synth.rot_y(0.7) @ synth.rot_x(-0.3)
    T[:3, 3] = [10.0, -4.0, 500.0]
    q = orc.transform_points(p, T)
    assert np.allclose(q, p @ T[:3, :3].T + T[:3, 3], rtol=0, atol=1e-10)
    assert np.allclose(orc.transform_points(q, np.linalg.inv(T)), p, rtol=0, atol=1e-9)


# ------------------------------------------------------------------ point-to-plane ICP (8f #4)
def test_icp_oracle_recovers_a_known_pose(orc):
    from defectproj import synth
    V, F = synth.param_mesh(30, 20, seed=4)
    V = V.astype(np.float64)
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    T = np.eye(4)
    T[:3, :3] = synth.rot_z(2.0) @ synth.rot_x(-1.5)
    T[:3, 3] = [0.8, -0.5, 0.6]
    src = orc.transform_points(V[::2], np.linalg.inv(T))
    Tr, fit, rmse, it, corr = orc.icp_point_to_plane(src, V, vn, 5.0)
    assert fit == 1.0 and rmse < 1e-9 and 1 <= it <= 30 and np.abs(Tr - T).max() < 1e-9
    assert np.array_equal(corr, np.arange(0, len(V), 2))              # every point found its own vertex
    T1, _, _, it1, _ = orc.icp_point_to_plane(src, V, vn, 5.0, max_iteration=1)
    assert it1 == 1 and np.abs(T1 - T).max() < np.abs(np.eye(4) - T).max()


# ------------------------------------------------------------------ normal estimation (8f #1 / #4)
def test_normals_oracle_against_numpy_eigh(orc):
    """The restated closed-form solver returns numpy's eigenvector of the smallest eigenvalue; the hybrid search
    keeps the max_nn nearest within the radius, the point itself included."""
    rng = np.random.default_rng(0)
    S = rng.normal(size=(600, 3))
    S /= np.linalg.norm(S, axis=1, keepdims=True)
    S *= 40.0
    N, cnt, covs = orc.estimate_normals(S, 12.0, 30)
    w, v = np.linalg.eigh(covs)
    assert np.abs(np.abs((v[:, :, 0] * N).sum(1)) - 1.0).max() < 1e-9
    assert np.abs((N * S).sum(1) / 40.0).min() > 0.97                 # radial on a sphere
    assert cnt.max() <= 30 and cnt.min() >= 3
    i = 17
    d2 = ((S - S[i]) ** 2).sum(1)
    assert cnt[i] == min(30, int((d2 < 144.0).sum()))
    assert orc.hybrid_neighbours(S, i, 12.0, 30)[0] == i
    flat = np.c_[rng.uniform(-1, 1, (200, 2)), np.zeros(200)]
    assert np.all(np.abs(orc.estimate_normals(flat, 0.5, 30)[0][:, 2]) > 1 - 1e-12)
    lonely, c1, _ = orc.estimate_normals(S, 1e-3, 30)
    assert np.all(c1 == 1) and np.all(lonely == [0.0, 0.0, 1.0])     # fewer than 3 neighbours
    kept = orc.estimate_normals(S, 12.0, 30, normals=-S)[0]
    assert np.all((kept * S).sum(1) < 0.0)                            # existing normals keep their side
    assert orc.fast_eigen3x3(np.zeros((3, 3))) == (0.0, 0.0, 0.0)
    assert orc.fast_eigen3x3(np.diag([3.0, 1.0, 2.0])) == (0.0, 1.0, 0.0)
