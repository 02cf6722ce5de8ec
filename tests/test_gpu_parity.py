"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the golden fixtures made from the reference, and -- at BASELINE.json's full sizes --
through size-independent properties.  Bars: bit-exact for integers / indices / face ids / t_hit (the float32
arithmetic contract of oracle.c semantic A); hit points within 1e-5 x bbox diagonal of the float64 truth on
rays that are not edge/grazing ties (north_star)."""
import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu

TOL_FRAC = 1e-5      # north_star: hit points within 1e-5 of the mesh bounding-box diagonal


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _grid_rays(orc, K, H, W, pose, step):
    from defectproj import Context
    ys, xs = np.mgrid[0:H:step, 0:W:step]
    xs, ys = xs.reshape(-1).astype(np.int64), ys.reshape(-1).astype(np.int64)
    return orc.rays_object_frame(xs, ys, Context.frame_xform(K, pose))


# ------------------------------------------------------------------------------------------ H1
@pytest.mark.parametrize("shape", [(1, 1), (3, 4), (5, 7), (96, 128), (97, 131), (2, 33, 65), (720, 1280), (3, 720, 1280)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_compaction_equals_numpy_where(ctx, shape, dt):
    rng = np.random.default_rng(hash(shape) % 1000)
    h = rng.random(shape).astype(dt)
    for thr in (0.5, 0.75, 0.0, 1.5):
        pix, I, counts = ctx.compact(h, thr)
        cmp_thr = np.float32(thr) if dt == np.float32 else thr
        ref = np.nonzero(h.reshape(-1) > cmp_thr)[0]
        assert np.array_equal(pix, ref.astype(np.uint32))                  # row-major order, strict '>'
        assert np.array_equal(I, h.reshape(-1)[ref].astype(np.float32))
        h3 = h.reshape((-1,) + h.shape[-2:])
        assert np.array_equal(counts, (h3 > cmp_thr).reshape(len(h3), -1).sum(1))


def test_compaction_edge_cases(ctx, golden):
    pix, I, c = ctx.compact(np.zeros((5, 7), np.float32), 0.5)
    assert len(pix) == 0 and c.tolist() == [0]
    pix, _, _ = ctx.compact(np.ones((1024, 1024), np.float32), 0.5)
    assert np.array_equal(pix, np.arange(1 << 20, dtype=np.uint32))
    h = np.array([[np.nan, 0.5, np.inf], [0.6, -np.inf, 0.5000001]], np.float64)
    pix, I, _ = ctx.compact(h, 0.5)
    assert pix.tolist() == [2, 3, 5]                                       # NaN and == thr never pass
    # values that differ from thr only beyond float32 precision must be decided in float64
    h = np.full((4, 4), 0.5, np.float64)
    h[2, 1] = np.nextafter(0.5, 1.0)
    pix, _, _ = ctx.compact(h, 0.5)
    assert pix.tolist() == [9]
    # the reference's own example and Gaussian (fixtures made by the reference's heatmap_to_points)
    for heat, thr, key in ((golden["g1_heat"], 0.5, "g1_points"), (golden["g2_heat"], 0.5, "g2_points_050"),
                           (golden["g2_heat"], 0.75, "g2_points_075")):
        pix, _, _ = ctx.compact(heat, thr)
        ref = golden[key]
        W = heat.shape[1]
        assert np.array_equal(pix, (ref[:, 1] * W + ref[:, 0]).astype(np.uint32))


def test_compaction_capacity_error(ctx, built_lib):
    import ctypes as C
    h = np.ones((8, 8), np.float32)
    pix = np.empty(10, np.uint32)
    n = C.c_int64(0)
    rc = built_lib.dp_compact(ctx._h, h.ctypes.data_as(C.c_void_p), 0, 1, 8, 8, 0.5, pix.ctypes.data_as(C.c_void_p),
                              None, 10, C.byref(n), None, 0, None)
    assert rc == -3 and n.value == 64 and b"capacity" in built_lib.dp_last_error(ctx._h)
    assert np.array_equal(pix, np.arange(10, dtype=np.uint32))             # the first `cap` entries are valid


# ------------------------------------------------------------------------------------------ H2
def test_compute_rays_bit_exact_vs_reference(ctx, golden):
    ref = golden["g2_points_050"]
    rays = ctx.compute_rays(ref[:, 0], ref[:, 1], golden["g4_K"])
    assert np.array_equal(rays, golden["g4_rays"])                         # float64, bit for bit


# ------------------------------------------------------------------------------------------ (3) structure
@pytest.mark.parametrize("n", [1, 2, 31, 32, 4095, 4096, 4097, 70001, 1 << 20])
def test_radix_sort_is_stable_argsort(ctx, n):
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << 30, n, dtype=np.uint32)
    if n > 100:
        k[rng.integers(0, n, n // 2)] = k[0]                               # many duplicates
    v = np.arange(n, dtype=np.uint32)
    ks, vs = ctx.radix_sort(k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(ks, k[order]) and np.array_equal(vs, v[order])


def _morton_ref(V, F):
    tri = V[F]                                                              # [nF,3,3] float32
    lo, hi = tri.min(1), tri.max(1)
    slo, shi = lo.min(0), hi.max(0)
    c = np.float32(0.5) * (lo + hi)
    ext = shi - slo
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (c - slo) * (np.float32(1024.0) / ext)
    g = np.where(ext > 0, g, np.float32(0))
    q = np.clip(g, 0, 1023).astype(np.uint32)

    def ex(v):
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    return (ex(q[:, 0]) << np.uint32(2)) | (ex(q[:, 1]) << np.uint32(1)) | ex(q[:, 2])


def test_morton_codes_match_numpy(ctx):
    V, F = synth.param_mesh(40, 25, seed=3)
    ctx.set_mesh(V, F).build_bvh()
    codes = ctx.morton_codes()
    assert codes.max() < (1 << 30)
    assert np.array_equal(codes, _morton_ref(V, F))


def _decode_nodes(nodes):
    """nodes uint32 [n,20] -> dict of per-node fields"""
    b = nodes.view(np.uint8).reshape(len(nodes), 80)
    origin = nodes[:, 0:3].copy().view(np.float32)
    e = b[:, 12:15].astype(np.int32) - 127
    imask = b[:, 15]
    child_base, tri_base = nodes[:, 4], nodes[:, 5]
    meta = b[:, 24:32]
    qlo = np.stack([b[:, 32:40], b[:, 40:48], b[:, 48:56]], axis=2).astype(np.float64)   # [n,8,3]
    qhi = np.stack([b[:, 56:64], b[:, 64:72], b[:, 72:80]], axis=2).astype(np.float64)
    scale = np.ldexp(1.0, e)[:, None, :]
    lo = origin[:, None, :].astype(np.float64) + qlo * scale
    hi = origin[:, None, :].astype(np.float64) + qhi * scale
    return dict(origin=origin, imask=imask, child_base=child_base, tri_base=tri_base, meta=meta, lo=lo, hi=hi)


def _check_structure(nodes, tris, V, F):
    d = _decode_nodes(nodes)
    n = len(nodes)
    face_of = tris[:, 3].copy().view(np.int32)
    assert sorted(face_of.tolist()) == list(range(len(F))), "every triangle in exactly one record"
    # records hold the vertices of their face
    assert np.array_equal(tris[:, 0:3], V[F[face_of, 0]]) and np.array_equal(tris[:, 4:7], V[F[face_of, 1]])
    assert np.array_equal(tris[:, 8:11], V[F[face_of, 2]])
    seen_tri = np.zeros(len(tris), np.int32)
    seen_node = np.zeros(n, np.int32)
    seen_node[0] = 1
    # exact box of every node = union of its children, computed bottom-up (children have larger indices)
    nlo = np.full((n, 3), np.inf)
    nhi = np.full((n, 3), -np.inf)
    for w in range(n - 1, -1, -1):
        k_inner = 0
        for s in range(8):
            m = int(d["meta"][w, s])
            if m == 0:
                assert not (d["imask"][w] >> s) & 1
                continue
            if (m & 0x18) == 0x18:
                assert (m >> 5) == 1 and (m & 7) == s and (d["imask"][w] >> s) & 1
                c = int(d["child_base"][w]) + k_inner
                k_inner += 1
                assert w < c < n
                seen_node[c] += 1
                clo, chi = nlo[c], nhi[c]
            else:
                assert not (d["imask"][w] >> s) & 1
                cnt = bin(m >> 5).count("1")
                assert (m >> 5) in (1, 3, 7)
                t0 = int(d["tri_base"][w]) + (m & 31)
                seen_tri[t0:t0 + cnt] += 1
                pts = tris[t0:t0 + cnt].reshape(cnt, 3, 4)[:, :, :3].reshape(-1, 3).astype(np.float64)
                clo, chi = pts.min(0), pts.max(0)
            # quantised child box must contain the exact child box
            assert (d["lo"][w, s] <= clo).all() and (d["hi"][w, s] >= chi).all(), (w, s)
            nlo[w] = np.minimum(nlo[w], clo)
            nhi[w] = np.maximum(nhi[w], chi)
        assert k_inner == bin(int(d["imask"][w])).count("1")
    assert (seen_tri == 1).all(), "every record referenced by exactly one leaf slot"
    assert (seen_node == 1).all(), "every node has exactly one parent"
    return nlo, nhi


@pytest.mark.parametrize("cfg", ["tiny", "small"])
def test_wide_bvh_structure(ctx, cfg):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[cfg], seed=5)
    ctx.set_mesh(V, F).build_bvh()
    nodes, tris = ctx.dump_bvh("object")
    nlo, nhi = _check_structure(nodes, tris, V, F)
    assert (nlo[0] <= V.min(0)).all() and (nhi[0] >= V.max(0)).all()
    # refit: same topology, boxes of the posed mesh
    T = synth.fixed_pose()
    ctx.pose_mesh(T)
    Vp = ctx.posed_vertices()
    nodes2, tris2 = ctx.dump_bvh("camera")
    assert np.array_equal(nodes2[:, 4:8], nodes[:, 4:8])                   # bases + meta unchanged
    _check_structure(nodes2, tris2, Vp, F)


@pytest.mark.parametrize("cfg", ["small", "c1_30k", "c2_500k"])
def test_single_launch_collapse_equals_per_level_launches(ctx, orc, cfg, monkeypatch):
    """The cooperative single-launch collapse (grid barriers) and the per-level launches (host read-back per level) build
    the same wide tree: same node and level counts, valid structure, identical hits."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[cfg], seed=5)
    K, H, W = synth.camera_720p()
    rays6 = _grid_rays(orc, K, H, W, synth.fixed_pose(), 6)
    got = []
    for mode in ("0", "1"):
        monkeypatch.setenv("DP_COLLAPSE_LAUNCHES", mode)
        ctx.set_mesh(V, F).build_bvh()
        st = ctx.stats()
        if cfg != "c2_500k":
            nodes, tris = ctx.dump_bvh("object")
            _check_structure(nodes, tris, V, F)
        t, f = ctx.cast_rays(rays6)
        got.append((st["n_wide_nodes"], st["wide_depth"], t, f))
    assert got[0][0] == got[1][0] and got[0][1] == got[1][1]
    assert np.array_equal(got[0][3], got[1][3]) and np.array_equal(_bits(got[0][2]), _bits(got[1][2]))
    assert (got[0][3] >= 0).sum() > 100


def test_duplicate_morton_codes_and_degenerate_triangles(ctx, orc):
    # 3000 triangles crammed into a few Morton cells + zero-area triangles: Karras must still build a tree
    rng = np.random.default_rng(9)
    base = rng.normal(size=(40, 3)).astype(np.float32)
    V = np.concatenate([base + np.float32(1e-4) * rng.normal(size=(40, 3)).astype(np.float32) for _ in range(75)])
    V[:, 2] += 6
    F = rng.integers(0, len(V), (3000, 3)).astype(np.int32)
    F[::50, 1] = F[::50, 0]                                                # degenerate
    ctx.set_mesh(V, F).build_bvh()
    nodes, tris = ctx.dump_bvh("object")
    _check_structure(nodes, tris, V, F)
    rays = np.zeros((4000, 6), np.float32)
    rays[:, 3:5] = rng.normal(size=(4000, 2)) * 0.3
    rays[:, 5] = 1
    t, f = ctx.cast_rays(rays)
    t0, f0 = orc.cast_brute_f32(V, F, rays)
    assert np.array_equal(f, f0) and np.array_equal(_bits(t), _bits(t0))


# ------------------------------------------------------------------------------------------ H4
@pytest.mark.parametrize("cfg,step", [("tiny", 4), ("small", 5), ("c1_30k", 7)])
def test_cast_rays_equals_brute_force(ctx, orc, cfg, step):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[cfg], seed=1)
    K, H, W = synth.camera_720p()
    rays6 = _grid_rays(orc, K, H, W, synth.fixed_pose(), step)
    ctx.set_mesh(V, F).build_bvh()
    t, f = ctx.cast_rays(rays6)
    t0, f0 = orc.cast_brute_f32(V, F, rays6)
    assert (f0 >= 0).sum() > 100
    assert np.array_equal(f, f0), f"{(f != f0).sum()} face ids differ"
    assert np.array_equal(_bits(t), _bits(t0))
    assert np.array_equal(np.isinf(t), f < 0)


def test_cast_rays_hard_cases(ctx, orc):
    """rays through vertices and along edges (exact ties), axis-parallel rays with zero components,
    origins inside the mesh, rays pointing away, zero-length directions."""
    V, F = synth.param_mesh(40, 25, seed=2)
    ctx.set_mesh(V, F).build_bvh()
    rng = np.random.default_rng(3)
    o = np.array([0.0, 0.0, 300.0], np.float32)
    parts = []
    d = V[rng.integers(0, len(V), 3000)] - o                              # straight at vertices
    parts.append(np.hstack([np.tile(o, (len(d), 1)), d]))
    mid = 0.5 * (V[F[:3000, 0]] + V[F[:3000, 1]]) - o                     # edge midpoints
    parts.append(np.hstack([np.tile(o, (len(mid), 1)), mid]))
    ax = np.zeros((600, 6), np.float32)                                   # axis-parallel, d has exact zeros
    ax[:, :3] = rng.uniform(-90, 90, (600, 3))
    ax[:200, 2] = 200; ax[:200, 5] = -1
    ax[200:400, 0] = -200; ax[200:400, 3] = 1
    ax[400:, 1] = 200; ax[400:, 4] = -1
    parts.append(ax)
    ins = np.zeros((2000, 6), np.float32)                                 # origins inside the tube
    ins[:, 0] = 60.0
    ins[:, 3:] = rng.normal(size=(2000, 3))
    parts.append(ins)
    away = np.hstack([np.tile(o, (100, 1)), np.tile(np.array([0, 0, 1], np.float32), (100, 1))])
    parts.append(away)
    zero = np.zeros((4, 6), np.float32)                                   # d = 0: defined as a miss
    parts.append(zero)
    rays6 = np.ascontiguousarray(np.vstack(parts), np.float32)
    t, f = ctx.cast_rays(rays6)
    t0, f0 = orc.cast_brute_f32(V, F, rays6)
    assert np.array_equal(f, f0), np.nonzero(f != f0)[0][:10]
    assert np.array_equal(_bits(t), _bits(t0))
    assert (f[-104:] == -1).all()


def test_cast_rays_incoherent_random(ctx, orc):
    rng = np.random.default_rng(5)
    V, F = synth.param_mesh(150, 100, seed=2)
    ctx.set_mesh(V, F).build_bvh()
    o = rng.normal(size=(30000, 3)).astype(np.float32) * 150
    d = -o + rng.normal(size=(30000, 3)).astype(np.float32) * 40
    rays6 = np.hstack([o, d]).astype(np.float32)                          # directions NOT normalised
    t, f = ctx.cast_rays(rays6)
    t0, f0 = orc.Bvh(V, F).cast_f32(rays6)
    assert np.array_equal(f, f0) and np.array_equal(_bits(t), _bits(t0))


def test_empty_and_tiny_meshes(ctx, orc):
    rng = np.random.default_rng(1)
    rays = np.zeros((300, 6), np.float32)
    rays[:, 3:5] = rng.normal(size=(300, 2)) * 0.3
    rays[:, 5] = 1
    ctx.set_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)).build_bvh()
    t, f = ctx.cast_rays(rays)
    assert np.isinf(t).all() and (f == -1).all()
    t, f = ctx.cast_rays(np.zeros((0, 6), np.float32))
    assert len(t) == 0
    for n in (1, 2, 3, 4, 5, 8, 9, 25):
        V = rng.normal(size=(3 * n, 3)).astype(np.float32)
        V[:, 2] += 5
        F = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
        ctx.set_mesh(V, F).build_bvh()
        t, f = ctx.cast_rays(rays)
        t0, f0 = orc.cast_brute_f32(V, F, rays)
        assert np.array_equal(f, f0) and np.array_equal(_bits(t), _bits(t0)), n


# ------------------------------------------------------------------------------------------ fused path
def _oracle_frame(orc, V, F, heat, thr, K, pose):
    from defectproj import Context
    xs, ys, I = orc.heatmap_to_points(heat, thr)
    rays6 = orc.rays_object_frame(xs, ys, Context.frame_xform(K, pose))
    t, f = orc.Bvh(V, F).cast_f32(rays6)
    return xs, ys, I, rays6, t, f


@pytest.mark.parametrize("hdt", [np.float32, np.float64])
def test_project_object_frame_c1(ctx, orc, hdt):
    """BASELINE configs[0]: 30k-triangle mesh, 720p Gaussian heatmap, thr 0.5 and 0.75, fixed pose."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    heat = synth.gaussian_heatmap((H, W), dtype=hdt)
    ctx.set_mesh(V, F).build_bvh()
    for thr, count in ((0.5, 10885), (0.75, 4501)):
        ctx.accum_reset()
        res = ctx.project(heat, K, pose[None], thr, "object", True,
                          want=("pixel", "intensity", "t_hit", "face", "point", "point64"))
        xs, ys, I, rays6, t, f = _oracle_frame(orc, V, F, heat, thr, K, pose)
        assert res["n"] == len(xs) == count
        assert np.array_equal(res["pixel"], (ys * W + xs).astype(np.uint32))
        assert np.array_equal(res["intensity"], I.astype(np.float32))
        assert np.array_equal(res["face"], f) and np.array_equal(_bits(res["t_hit"]), _bits(t))
        hit = f >= 0
        assert res["hits"] == hit.sum() > 1000
        d = orc.compute_rays(xs, ys, K)
        pref = d[hit] * t[hit].astype(np.float64)[:, None]               # :261-263, origin 0
        assert np.array_equal(res["point64"][hit], pref)
        assert np.isnan(res["point64"][~hit]).all() and np.isnan(res["point"][~hit]).all()
        assert np.array_equal(res["point"][hit], pref.astype(np.float32))
        hist, fmax, vmax = ctx.accum_get()
        h0, f0, v0 = orc.accumulate(f, I.astype(np.float32), F, len(V))
        assert np.array_equal(hist, h0) and np.array_equal(fmax, f0) and np.array_equal(vmax, v0)
        assert hist.sum() == res["hits"]


def test_project_against_float64_truth_north_star_criterion(ctx, orc):
    """face ids bit-exact except rays the float64 classifier flags as edge/grazing ties; hit points
    within 1e-5 x bbox diagonal."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    heat = synth.blob_heatmap((H, W), seed=4)
    ctx.set_mesh(V, F).build_bvh()
    res = ctx.project(heat, K, pose[None], 0.3, "object", False, want=("pixel", "t_hit", "face", "point64"))
    xs, ys, I, rays6, _, _ = _oracle_frame(orc, V, F, heat, 0.3, K, pose)
    bvh = orc.Bvh(V, F)
    t64, f64, tie = bvh.cast_f64(rays6)
    clean = tie == 0
    assert clean.mean() > 0.95
    assert np.array_equal(res["face"][clean], f64[clean])
    hit = clean & (f64 >= 0)
    assert hit.sum() > 1000
    diag = float(np.linalg.norm(V.max(0) - V.min(0)))
    d = orc.compute_rays(xs, ys, K)
    p64 = d[hit] * t64[hit][:, None]
    assert np.linalg.norm(res["point64"][hit] - p64, axis=1).max() <= TOL_FRAC * diag
    # on tie rays the GPU's face must still be a genuine candidate: it lies within the tie margin
    tie_hit = (~clean) & (res["face"] >= 0)
    if tie_hit.any():
        tt, margin, _ = bvh.eval_face(rays6[tie_hit], res["face"][tie_hit])
        tau, tau_t = bvh.margins(rays6)
        assert (margin >= -tau).all()
        ok = (f64[tie_hit] < 0) | (np.abs(tt - np.where(f64[tie_hit] >= 0, t64[tie_hit], tt)) <= 4 * tau_t)
        assert ok.all()


def test_project_camera_frame_is_reference_literal(ctx, orc):
    """DP_FRAME_CAMERA: vertices posed in float64 then cast (:549-550, :245), BVH refitted, rays from (0,0,0)."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    heat = synth.gaussian_heatmap((H, W), dtype=np.float64)
    ctx.set_mesh(V.astype(np.float64), F).build_bvh()
    ctx.pose_mesh(pose)
    Vref = orc.pose_vertices(V.astype(np.float64), pose)
    assert np.array_equal(ctx.posed_vertices(), Vref)
    ctx.accum_reset()
    res = ctx.project(heat, K, None, 0.5, "camera", True, want=("pixel", "t_hit", "face", "point64"))
    xs, ys, I = orc.heatmap_to_points(heat, 0.5)
    d = orc.compute_rays(xs, ys, K)
    t, f = orc.Bvh(Vref, F).cast_f32(orc.rays6_camera(d))
    assert np.array_equal(res["face"], f) and np.array_equal(_bits(res["t_hit"]), _bits(t))
    hist, fmax, vmax = ctx.accum_get()
    h0, f0, v0 = orc.accumulate(f, I.astype(np.float32), F, len(V))
    assert np.array_equal(hist, h0) and np.array_equal(fmax, f0) and np.array_equal(vmax, v0)
    # object-frame and camera-frame agree wherever neither is a tie (rigid invariance)
    res_o = ctx.project(heat, K, pose[None], 0.5, "object", False, want=("face", "point64"))
    rays6 = orc.rays_object_frame(xs, ys, orc.frame_xform(K, pose))
    _, _, tie = orc.Bvh(V, F).cast_f64(rays6)
    clean = tie == 0
    assert np.array_equal(res_o["face"][clean], res["face"][clean])
    both = clean & (f >= 0)
    diag = float(np.linalg.norm(V.max(0) - V.min(0)))
    assert np.linalg.norm(res_o["point64"][both] - res["point64"][both], axis=1).max() <= TOL_FRAC * diag


def test_project_batch_of_frames_single_launch(ctx, orc):
    """config 3 shape at test size: several views, per-frame pose, one launch, histogram accumulated over frames."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["small"], seed=2)
    K, H, W = synth.K_matrix(150.0, 150.0, 80.0, 60.0), 120, 160
    poses = synth.fibonacci_poses(6, radius=400.0)
    heats = np.stack([synth.blob_heatmap((H, W), seed=i) for i in range(6)])
    ctx.set_mesh(V, F).build_bvh()
    ctx.accum_reset()
    res = ctx.project(heats, K, poses, 0.4, "object", True, want=("pixel", "intensity", "t_hit", "face"))
    bvh = orc.Bvh(V, F)
    faces, ts, pixs, Is = [], [], [], []
    for b in range(6):
        xs, ys, I = orc.heatmap_to_points(heats[b], 0.4)
        t, f = bvh.cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(K, poses[b])))
        faces.append(f); ts.append(t); Is.append(I); pixs.append(b * H * W + ys * W + xs)
    f_all, t_all, I_all = np.concatenate(faces), np.concatenate(ts), np.concatenate(Is)
    assert np.array_equal(res["pixel"], np.concatenate(pixs).astype(np.uint32))
    assert np.array_equal(res["face"], f_all) and np.array_equal(_bits(res["t_hit"]), _bits(t_all))
    hist, fmax, vmax = ctx.accum_get()
    h0, f0, v0 = orc.accumulate(f_all, I_all, F, len(V))
    assert (f_all >= 0).sum() > 500
    assert np.array_equal(hist, h0) and np.array_equal(fmax, f0) and np.array_equal(vmax, v0)


@pytest.mark.parametrize("kind", ["dense", "sparse", "batch", "camera"])
def test_in_kernel_rays_and_points_equal_the_separate_kernels(ctx, kind, monkeypatch):
    """dp_project generates the rays and writes the hit points inside the traversal kernel (default); DP_FUSE_RAYS=0 runs
    the separate k_raygen / k_points kernels around it.  Same operations in the same order: every output bit-identical,
    for the packet traversal (dense), the eight-lanes-per-ray traversal (sparse), a batch of frames and the camera frame;
    both node sets."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0, scale=6.0)
    pose = synth.fill_frame_pose()
    ctx.set_mesh(V.astype(np.float64), F).build_bvh()
    H, W = 384, 512
    K = synth.K_matrix(126.0, 126.0, W / 2, H / 2)
    if kind == "dense":
        heat, poses, frame = np.ones((H, W), np.float32), pose[None], "object"
    elif kind == "sparse":
        heat, poses, frame = synth.blob_heatmap((H, W), seed=9), pose[None], "object"
    elif kind == "batch":
        heat = np.stack([synth.blob_heatmap((H, W), seed=20 + i, dtype=np.float64) for i in range(3)])
        poses, frame = np.stack([pose, pose @ synth._pose(synth.rot_z(7.0), [3.0, 1.0, -2.0]), pose]), "object"
    else:
        heat, poses, frame = np.ones((H, W), np.float32), None, "camera"
        ctx.pose_mesh(pose)
    want = ("pixel", "intensity", "t_hit", "face", "point", "point64")
    got = {}
    for fat in ("1", "0"):
        monkeypatch.setenv("DP_FAT", fat)
        for fuse in ("1", "0"):
            monkeypatch.setenv("DP_FUSE_RAYS", fuse)
            ctx.accum_reset()
            r = ctx.project(heat, K, poses, 0.5, frame, True, want=want)
            r["acc"] = ctx.accum_get()
            got[(fat, fuse)] = r
    ref = got[("1", "0")]
    assert ref["hits"] > 1000
    for key, r in got.items():
        assert r["n"] == ref["n"] and r["hits"] == ref["hits"], key
        for k in want:
            a, b = r[k], ref[k]
            assert np.array_equal(a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64),
                                  b.view(np.uint32 if b.dtype.itemsize == 4 else np.uint64)), (key, k)
        for a, b in zip(r["acc"], ref["acc"]):
            assert np.array_equal(a, b), key


def test_project_empty_selection_and_miss_all(ctx):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["tiny"], seed=0)
    K, H, W = synth.camera_720p()
    ctx.set_mesh(V, F).build_bvh()
    res = ctx.project(np.zeros((H, W), np.float32), K, synth.fixed_pose()[None], 0.5, "object", True)
    assert res["n"] == 0 and res["hits"] == 0 and len(res["face"]) == 0
    far = synth.fixed_pose()
    far[:3, 3] = [5000.0, 0, 600.0]
    res = ctx.project(synth.gaussian_heatmap((H, W), dtype=np.float32), K, far[None], 0.9, "object", True)
    assert res["n"] > 0 and res["hits"] == 0 and (res["face"] == -1).all() and np.isinf(res["t_hit"]).all()


def test_error_behaviour(built_lib):
    from defectproj import Context, DefectProjError
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["tiny"], seed=0)
    K, H, W = synth.camera_720p()
    with Context(0) as c:
        with pytest.raises(DefectProjError, match="no mesh"):
            c.build_bvh()
        c.set_mesh(V, F)
        with pytest.raises(DefectProjError, match="no BVH"):
            c.cast_rays(np.zeros((1, 6), np.float32))
        with pytest.raises(DefectProjError, match="no BVH"):
            c.project(np.ones((4, 4), np.float32), K, np.eye(4)[None], 0.5)
        c.build_bvh()
        with pytest.raises(DefectProjError, match="dp_pose_mesh"):
            c.project(np.ones((4, 4), np.float32), K, None, 0.5, frame="camera")
        with pytest.raises(ValueError):
            c.project(np.ones((4, 4), np.float32), K, np.eye(4)[None].repeat(2, 0), 0.5)
        bad = F.copy()
        bad[3, 1] = len(V)
        with pytest.raises(ValueError, match="out of range"):
            c.set_mesh(V, bad)
        import torch
        with pytest.raises(ValueError, match="out of range"):              # device-resident meshes are checked too
            c.set_mesh(torch.from_numpy(V).cuda(), torch.from_numpy(bad).cuda())
        c.set_mesh(torch.from_numpy(V).cuda(), torch.from_numpy(F).cuda()).build_bvh()
        with pytest.raises(ValueError):
            c.pose_mesh(np.ones((4, 4)))
    with pytest.raises(ValueError, match="out of range"):
        Context(99)


# ------------------------------------------------------------------------------------------ façade vs golden
def test_facade_matches_reference_run(golden, tmp_path, built_lib):
    """The reference's own ray_tracing() output (tests/golden/make_golden.py) reproduced through the drop-in."""
    from defectproj import defect_projection as dpj
    K = golden["g5_K"]
    synth.write_scene_dir(str(tmp_path), K, (96, 128), color_to_depth=golden["g5_color_to_depth"])
    mesh = dpj.TriangleMesh(golden["g5_V_depthcam"], golden["g5_F"])
    for thr, tag in ((0.5, "050"), (0.75, "075")):
        pcd, mesh_out = dpj.ray_tracing(str(tmp_path), mesh, golden["g2_heat"], K, heatmap_threshold=thr)
        assert isinstance(pcd, dpj.PointCloud)
        assert np.array_equal(np.asarray(pcd.points), golden[f"g5_points_{tag}"])
        assert np.allclose(np.asarray(pcd.colors), golden[f"g5_colors_{tag}"], rtol=0, atol=1e-12)
        # float64 posing: same values up to the summation order of the 4x4 product (numpy/BLAS in the fixture)
        assert np.allclose(np.asarray(mesh_out.vertices), golden[f"g5_V_colorcam_{tag}"], rtol=1e-14, atol=1e-11)
        assert np.array_equal(np.asarray(mesh_out.vertices).astype(np.float32),
                              golden[f"g5_V_colorcam_{tag}"].astype(np.float32))
        hist, fmax, vmax = dpj.face_intensities()
        assert hist.sum() == len(pcd.points) and np.array_equal(np.bincount(pcd.face_ids, minlength=len(hist)), hist)
    # step-by-step surface
    pts = dpj.heatmap_to_points(golden["g2_heat"], 0.5)
    assert np.array_equal(np.array([[p[0], p[1], p[2]] for p in pts], np.float64), golden["g2_points_050"])
    rays, inten = dpj.compute_rays(pts, K)
    assert np.array_equal(rays, golden["g4_rays"]) and np.array_equal(inten, golden["g4_intensities"])
    rays2, _ = dpj.compute_rays(list(pts), K)                             # plain list of tuples, as the reference passes
    assert np.array_equal(rays2, rays)
    posed = dpj.TriangleMesh(golden["g5_V_colorcam_050"], golden["g5_F"])
    P, I = dpj.intersect_rays_with_mesh(posed, rays, np.array([0, 0, 0]), inten)
    assert np.array_equal(P, golden["g5_points_050"])
    # miss-all branch -> LineSet (:561-563)
    far = dpj.TriangleMesh(golden["g5_V_model"].astype(np.float64) + np.array([5000.0, 0, 0]), golden["g5_F"])
    ls, _ = dpj.ray_tracing(str(tmp_path), far, golden["g2_heat"], K, heatmap_threshold=0.9)
    assert isinstance(ls, dpj.LineSet)
    assert np.allclose(ls.points, golden["g6_lineset_points"], rtol=0, atol=1e-9)
    assert np.array_equal(ls.lines, golden["g6_lineset_lines"])
    # documented deviation: empty selection returns empty outputs instead of raising
    ls, _ = dpj.ray_tracing(str(tmp_path), mesh, np.zeros((96, 128)), K)
    assert len(ls.lines) == 0
    assert dpj.heatmap_to_points(np.zeros((4, 4)), 0.5) == []


# ------------------------------------------------------------------------------------------ full size
@pytest.mark.parametrize("cfg", ["c2_500k", "ns_1m", "c4_5m"])
def test_full_size_dense_frame(ctx, orc, cfg):
    """BASELINE configs[1], the north-star 1M-triangle mesh and configs[3] (5M triangles, a hierarchy larger than L2,
    which also switches the traversal's L1 prefetch on): 1024x1024 dense frame, every ray checked
    against the oracle's own BVH caster, plus the size-independent properties."""
    import torch
    V, F = synth.param_mesh(*synth.MESH_CONFIGS[cfg], seed=0, scale=6.0)
    K, H, W = synth.camera_wfov()
    pose = synth.fill_frame_pose()
    ctx.set_mesh(V, F).build_bvh()
    st = ctx.stats()
    assert st["n_tris"] == len(F) and st["n_wide_nodes"] < len(F) // 2
    heat = torch.rand((1, H, W), device="cuda") * 0.5 + 0.5             # all > 0.5 except exact 0.5
    heat[0, 5, 7] = 0.25
    n = H * W
    out = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), t_hit=torch.empty(n, device="cuda"),
               face=torch.empty(n, dtype=torch.int32, device="cuda"), intensity=torch.empty(n, device="cuda"))
    ctx.accum_reset()
    nr, nh = ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
    hnp = heat.cpu().numpy()[0]
    xs, ys, I = orc.heatmap_to_points(hnp, 0.5)
    assert nr == len(xs) >= n - 100
    t, f = orc.Bvh(V, F).cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(K, pose)))
    gf, gt = out["face"][:nr].cpu().numpy(), out["t_hit"][:nr].cpu().numpy()
    assert np.array_equal(out["pixel"][:nr].cpu().numpy().view(np.uint32), (ys * W + xs).astype(np.uint32))
    assert np.array_equal(gf, f), f"{(gf != f).sum()} of {nr} face ids differ"
    assert np.array_equal(_bits(gt), _bits(t))
    hist, fmax, vmax = ctx.accum_get()
    assert hist.sum() == nh == (f >= 0).sum() and nh > 0.9 * nr          # fill-frame pose: >= 90 % hit
    assert np.array_equal(hist, np.bincount(f[f >= 0], minlength=len(F)).astype(np.int32))
    assert fmax.max() <= I.max() and vmax.max() == fmax.max()
    assert ((fmax > 0) == (hist > 0)).all()
    # idempotence: a second accumulation doubles the histogram and leaves the maxima alone
    ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
    h2, f2, v2 = ctx.accum_get()
    assert np.array_equal(h2, 2 * hist) and np.array_equal(f2, fmax) and np.array_equal(v2, vmax)


def test_shards_of_a_batch_combine_to_the_whole(ctx):
    """Multi-GPU semantics on one GPU: frames split into rank blocks (shard_range), each block projected on its
    own, integer histograms summed and maxima maxed == the single-launch result, bit for bit."""
    from defectproj.projector import shard_range
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    B = 8
    poses = synth.helix_poses(B, turns=1)
    heats = np.stack([synth.blob_heatmap((H, W), seed=100 + i) for i in range(B)])
    ctx.set_mesh(V, F).build_bvh()
    ctx.accum_reset()
    whole = ctx.project(heats, K, poses, 0.5, "object", True, want=("pixel", "face", "t_hit"))
    hw, fw, vw = ctx.accum_get()
    for world in (2, 4, 8):
        hs, fs, vs, faces, pix = [], [], [], [], []
        for r in range(world):
            lo, hi = shard_range(B, world, r)
            ctx.accum_reset()
            part = ctx.project(heats[lo:hi], K, poses[lo:hi], 0.5, "object", True, want=("pixel", "face"))
            a, b, c = ctx.accum_get()
            hs.append(a); fs.append(b); vs.append(c); faces.append(part["face"])
            pix.append(part["pixel"].astype(np.int64) + lo * H * W)
        assert np.array_equal(np.sum(hs, axis=0, dtype=np.int32), hw)
        assert np.array_equal(np.max(fs, axis=0), fw) and np.array_equal(np.max(vs, axis=0), vw)
        assert np.array_equal(np.concatenate(faces), whole["face"])
        assert np.array_equal(np.concatenate(pix), whole["pixel"].astype(np.int64))


def test_repeated_calls_are_identical_and_schedule_state_is_safe(ctx, orc):
    """The traversal learns a packet schedule from the previous launch over the same rays (heavy packets
    first).  It must never change results, and must be dropped when the ray count changes."""
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0, scale=6.0)
    K, H, W = synth.camera_wfov()
    pose = synth.fill_frame_pose()
    ctx.set_mesh(V, F).build_bvh()
    dense = np.ones((256, 512), np.float32)
    Ks = synth.K_matrix(126.0, 126.0, 256.0, 128.0)
    ref = None
    for it in range(4):                                              # 1st: natural order, 2nd: thresholds, 3rd+: lists
        ctx.accum_reset()
        res = ctx.project(dense, Ks, pose[None], 0.5, "object", True, want=("pixel", "t_hit", "face"))
        hist = ctx.accum_get()[0]
        if ref is None:
            xs, ys, I = orc.heatmap_to_points(dense, 0.5)
            t, f = orc.Bvh(V, F).cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(Ks, pose)))
            assert np.array_equal(res["face"], f) and np.array_equal(_bits(res["t_hit"]), _bits(t))
            ref = (res["face"].copy(), res["t_hit"].copy(), hist.copy())
        assert np.array_equal(res["face"], ref[0]) and np.array_equal(_bits(res["t_hit"]), _bits(ref[1]))
        assert np.array_equal(hist, ref[2]) and hist.sum() == res["hits"]
    # a different ray count right after: the learnt lists must be ignored, not indexed out of range
    for shape, thr in (((100, 300), 0.5), ((256, 512), 0.5), ((31, 33), 0.5)):
        heat = np.random.default_rng(shape[0]).random(shape).astype(np.float32)
        Kx = synth.K_matrix(126.0, 126.0, shape[1] / 2, shape[0] / 2)
        res = ctx.project(heat, Kx, pose[None], thr, "object", False, want=("pixel", "t_hit", "face"))
        xs, ys, I = orc.heatmap_to_points(heat, thr)
        t, f = orc.Bvh(V, F).cast_f32(orc.rays_object_frame(xs, ys, orc.frame_xform(Kx, pose)))
        assert np.array_equal(res["face"], f) and np.array_equal(_bits(res["t_hit"]), _bits(t))


def test_vertex_maxima_are_derived_on_read(ctx, orc):
    """vmax is not accumulated per hit: dp_accum_flush derives it from fmax (max over incident faces), which must
    equal the oracle's per-hit accumulation bit for bit, also through the raw device pointers."""
    import torch
    V, F = synth.param_mesh(40, 25, seed=7)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    ctx.set_mesh(V, F).build_bvh()
    ctx.accum_reset()
    heat = synth.blob_heatmap((H, W), seed=3)
    res = ctx.project(heat, K, pose[None], 0.3, frame="object", accumulate=True, want=("face", "intensity"))
    _, _, v_ref = orc.accumulate(res["face"], res["intensity"], F, len(V))
    hp, fp, vp = ctx.accum_device_ptrs()
    from defectproj.projector import _DevView
    v_dev = torch.as_tensor(_DevView(vp, len(V), "<f4"), device="cuda")
    assert float(v_dev.max()) == 0.0                       # nothing derived yet: the kernel only writes hist / fmax
    ctx.accum_flush(torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(v_dev.cpu().numpy(), v_ref) and v_ref.max() > 0
    assert np.array_equal(ctx.accum_get()[2], v_ref)       # idempotent


def test_statistics_counters(ctx):
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["small"], seed=0)
    K, H, W = synth.camera_720p()
    ctx.set_mesh(V, F).build_bvh()
    ctx.set_stats(True)
    try:
        res = ctx.project(synth.gaussian_heatmap((H, W), dtype=np.float32), K, synth.fixed_pose()[None], 0.5)
        st = ctx.stats()
        assert st["rays"] == res["n"] == 10885 and st["hits"] == res["hits"]
        assert st["nodes_fetched"] >= st["rays"] and st["tris_tested"] >= st["hits"]
        counts = ctx.ray_node_counts(res["n"])
        assert counts.sum() == st["nodes_fetched"] and counts.min() >= 1
        assert st["n_tris"] == len(F) and 1 <= st["wide_depth"] <= 16
    finally:
        ctx.set_stats(False)


def test_stage_timings_are_opt_in(ctx):
    """dp_project records its stage events only after dp_set_timing(1): five event records cost ~14 us per call."""
    from defectproj import DefectProjError
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["small"], seed=0)
    K, H, W = synth.camera_720p()
    ctx.set_mesh(V, F).build_bvh()
    heat = synth.gaussian_heatmap((H, W), dtype=np.float32)
    ctx.set_timing(False)
    ctx.project(heat, K, synth.fixed_pose()[None], 0.5)
    with pytest.raises(DefectProjError):
        ctx.last_timings()
    ctx.set_timing(True)
    try:
        ctx.project(heat, K, synth.fixed_pose()[None], 0.5)
        t = ctx.last_timings()
        assert 0 < t["trace_ms"] <= t["total_ms"] and t["compact_ms"] > 0 and t["raygen_ms"] > 0
    finally:
        ctx.set_timing(False)


@pytest.mark.parametrize("speculate", [True, False])
def test_frame_stream_equals_blocking_calls(ctx, orc, speculate):
    """The pipelined host-buffer API (H2D | kernels | D2H on three streams) returns what Context.project returns -- with
    the read-back queued behind the kernels on a prediction of the frame's size (sparse, empty and dense frames in turn:
    every misprediction path) and with the exact-size read-back after a host wait."""
    import torch
    from defectproj import FrameStream
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.K_matrix(305.0, 305.0, 320.0, 180.0), 360, 640
    B = 11
    poses = synth.helix_poses(B, turns=1)
    heats = [torch.from_numpy(synth.blob_heatmap((H, W), seed=40 + i)).pin_memory() for i in range(B)]
    heats[3] = torch.zeros((H, W)).pin_memory()                                   # an empty frame in the middle
    heats[5] = torch.ones((H, W)).pin_memory()                                    # a dense one: its pixel list is the identity
    heats[8] = torch.ones((H, W)).pin_memory()                                    # dense, dense, then sparse again
    heats[9] = torch.ones((H, W)).pin_memory()
    ctx.set_mesh(V, F).build_bvh()
    ctx.accum_reset()
    fs = FrameStream(ctx, H, W, want=("pixel", "t_hit", "face", "point"))
    fs.speculate = speculate
    got = {}
    for i, res in fs.run(heats, K, poses, 0.5, "object", True):
        got[i] = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in res.items()}
    hist_stream = ctx.accum_get()[0]
    assert sorted(got) == list(range(B))
    ctx.accum_reset()
    for i in range(B):
        ref = ctx.project(heats[i].numpy(), K, poses[i][None], 0.5, "object", True, want=("pixel", "t_hit", "face", "point"))
        assert got[i]["n"] == ref["n"] and got[i]["hits"] == ref["hits"]
        for k in ("pixel", "t_hit", "face", "point"):
            assert np.array_equal(got[i][k], ref[k], equal_nan=(k in ("t_hit", "point"))), (i, k)
    assert got[3]["n"] == 0
    assert got[5]["n"] == H * W and np.array_equal(got[5]["pixel"], np.arange(H * W, dtype=np.uint32))
    assert np.array_equal(hist_stream, ctx.accum_get()[0]) and hist_stream.sum() == sum(g["hits"] for g in got.values())


def test_frame_stream_modes_ship_only_what_was_asked(ctx):
    """FrameStream(want=...): 'payload' (face + point, no t_hit), accumulate-only (nothing per ray: counts and the device
    accumulators are the product).  Same hits, same histogram; only the bytes that cross PCIe differ."""
    import torch
    from defectproj import FrameStream
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.K_matrix(305.0, 305.0, 320.0, 180.0), 360, 640
    B = 4
    poses = synth.helix_poses(B, turns=1)
    heats = [torch.from_numpy(synth.blob_heatmap((H, W), seed=60 + i)).pin_memory() for i in range(B)]
    ctx.set_mesh(V, F).build_bvh()
    ref, hist = {}, {}
    for name, want in (("full", ("pixel", "t_hit", "face", "point")), ("payload", ("pixel", "face", "point")), ("acc", ())):
        ctx.accum_reset()
        fs = FrameStream(ctx, H, W, want=want)
        got = {i: {k: (v.copy() if hasattr(v, "copy") else v) for k, v in res.items()}
               for i, res in fs.run(heats, K, poses, 0.5, "object", True)}
        hist[name] = ctx.accum_get()[0]
        ref[name] = got
        assert fs.last_d2h_bytes == 16 + sum({"pixel": 4, "t_hit": 4, "face": 4, "point": 12}[k] for k in want) * got[B - 1]["n"]
    for i in range(B):
        assert ref["payload"][i]["n"] == ref["full"][i]["n"] == ref["acc"][i]["n"] > 0
        assert ref["payload"][i]["hits"] == ref["full"][i]["hits"] == ref["acc"][i]["hits"]
        assert np.array_equal(ref["payload"][i]["face"], ref["full"][i]["face"])
        assert np.array_equal(ref["payload"][i]["point"], ref["full"][i]["point"], equal_nan=True)
        assert "t_hit" not in ref["payload"][i] and set(ref["acc"][i]) == {"n", "hits"}
    assert np.array_equal(hist["payload"], hist["full"]) and np.array_equal(hist["acc"], hist["full"])
    assert hist["full"].sum() == sum(ref["full"][i]["hits"] for i in range(B)) > 0


# ------------------------------------------------------------------------------------------ depth path (8f #1)
def test_depth_projection_path_matches_reference(built_lib, orc):
    """heatmap_to_point3d / calc_coordinates / align_to_surface / depth_projection_heatmap through the drop-in
    against the outputs of the reference's own functions (tests/golden/depth_path.npz), bit for bit (float64)."""
    import os
    from defectproj import defect_projection as dpj
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_path.npz"))
    K, heat, depth = g["d_K"], g["d_heat"], g["d_depth"]
    for thr, key in ((0.1, "d_point3d_010"), (0.5, "d_point3d_050")):
        assert np.array_equal(dpj.heatmap_to_point3d(heat, depth, K, threshold=thr), g[key])
    assert np.array_equal(dpj.heatmap_to_point3d(heat, depth[:50, :70], K, threshold=0.3), g["d_point3d_small"])
    # float32 map: the reference divides in float32 and stores the rounded quotient (fixture from its own function)
    for thr, key in ((0.1, "d_point3d_f32_010"), (0.5, "d_point3d_f32_050")):
        assert np.array_equal(dpj.heatmap_to_point3d(heat.astype(np.float32), depth, K, threshold=thr), g[key])
    assert np.array_equal(dpj.heatmap_to_point3d(heat.astype(np.float32), depth, K, threshold=0.5),
                          orc.heatmap_to_point3d(heat.astype(np.float32), depth, K, 0.5))
    assert dpj.heatmap_to_point3d(np.zeros((8, 8)) + 1e-3, np.zeros((8, 8), np.uint16), K).shape == (0,)
    got = dpj.calc_coordinates(depth, [tuple(p) for p in g["d_picks"]], K)
    assert np.array_equal(got, g["d_calc"])
    target = dpj.PointCloud(g["d_target_points"], normals=g["d_target_normals"])
    offs, ali, p3 = dpj.depth_projection_heatmap(depth, K, target, heat)
    assert np.array_equal(p3, g["d_proj_point3d"])
    assert np.array_equal(ali, g["d_proj_aligned"]) and np.array_equal(offs, g["d_proj_offset"])
    o2, a2 = dpj.align_to_surface(g["d_point3d_050"], target, offset=0.1)
    assert np.array_equal(o2, g["d_align_offset_01"]) and np.array_equal(a2, g["d_align_aligned_01"])
    bare = dpj.PointCloud(g["d_target_points"])                   # no normals: estimated in place like :431-436
    o3, a3 = dpj.align_to_surface(g["d_point3d_050"], bare, offset=0.1)
    assert bare.has_normals() and np.array_equal(a3, a2)
    o4, _ = dpj.align_to_surface(g["d_point3d_050"], dpj.PointCloud(bare.points, normals=bare.normals), offset=0.1)
    assert np.array_equal(o3, o4)
    with pytest.raises(ValueError, match="uint16"):
        dpj.heatmap_to_point3d(heat, depth.astype(np.float32), K)


def test_depth_path_full_size_properties(ctx, orc):
    """720p heatmap + depth image, 200k-point target cloud: selection == numpy, every nearest neighbour exact."""
    rng = np.random.default_rng(11)
    K, H, W = synth.camera_720p()
    heat = synth.blob_heatmap((H, W), seed=2, dtype=np.float64)
    depth = rng.integers(400, 900, (H, W)).astype(np.uint16)
    depth[rng.random((H, W)) < 0.1] = 0
    pts = ctx.depth_backproject(heat, depth, K, 0.4)
    ref = orc.heatmap_to_point3d(heat, depth, K, 0.4)
    assert len(pts) > 50000 and np.array_equal(pts, ref)
    V, _ = synth.param_mesh(500, 400, seed=1)
    tp = V.astype(np.float64) * 3.0 + np.array([0.0, 0.0, 650.0])
    q = pts[:: max(1, len(pts) // 4000)]
    _, ali, idx = ctx.align_to_surface(q, tp, None, 0.0)
    ridx = orc.nearest_points(q, tp)
    assert np.array_equal(idx, ridx) and np.array_equal(ali, tp[ridx])


def test_align_to_surface_grid_search_equals_the_scan(ctx, monkeypatch):
    """dp_align_to_surface on large inputs walks a uniform grid over the target ring by ring; DP_NN_GRID=0 keeps the
    tiled scan.  Same indices, aligned and offset points bit for bit -- for queries on the surface, in the hollow of
    the torus, far outside the target's box, on duplicated target points (ties to the smaller index) and NaN."""
    rng = np.random.default_rng(5)
    V, F = synth.param_mesh(300, 200, seed=3)
    tp = V.astype(np.float64)
    tp = np.concatenate([tp, tp[:1000]])                             # duplicates: equal distances
    tn = rng.normal(size=tp.shape)
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    q = np.concatenate([tp[::9] + rng.normal(scale=0.4, size=tp[::9].shape),      # near the surface
                        tp[:1000],                                                # exactly on duplicated points
                        rng.uniform(-15.0, 15.0, (500, 3)),                       # the hollow around the axis
                        rng.uniform(-1.0, 1.0, (500, 3)) * 5000.0,                # far outside the box
                        np.zeros((1, 3))])
    q = np.c_[q, np.ones(len(q))]                                    # [x, y, z, intensity] like heatmap_to_point3d
    q[17, :3] = np.nan
    assert len(q) * len(tp) >= 2 ** 27
    monkeypatch.setenv("DP_NN_GRID", "1")
    o1, a1, i1 = ctx.align_to_surface(q, tp, tn, 0.5)
    monkeypatch.setenv("DP_NN_GRID", "0")
    o0, a0, i0 = ctx.align_to_surface(q, tp, tn, 0.5)
    assert np.array_equal(i1, i0) and i1[17] == -1 and (np.delete(i1, 17) >= 0).all()
    ok = i1 >= 0
    assert np.array_equal(a1[ok], a0[ok]) and np.array_equal(o1[ok], o0[ok])
    assert np.array_equal(a1[ok], tp[i1[ok]])
    on_dup = slice(len(tp[::9]), len(tp[::9]) + 1000)
    assert np.array_equal(i1[on_dup], np.arange(1000))               # the smaller of the two equal indices
    # without normals only the nearest points come back
    monkeypatch.setenv("DP_NN_GRID", "1")
    _, a2, i2 = ctx.align_to_surface(q, tp, None, 0.0)
    assert np.array_equal(i2, i1) and np.array_equal(a2[ok], a1[ok])


def _moved(V, rz, t):
    R = synth.rot_z(rz) @ synth.rot_x(0.4 * rz)
    return V.astype(np.float64) @ R.T + np.asarray(t, dtype=np.float64)


def test_update_vertices_refit_equals_a_rebuild(ctx):
    """dp_update_vertices keeps the hierarchy's topology and refits it to the new vertex positions; the closest hits
    do not depend on the hierarchy, so both frames give the bits of a fresh build on the moved mesh."""
    V, F = synth.param_mesh(60, 40, seed=11)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    heat = synth.gaussian_heatmap((H, W), dtype=np.float32)
    Tc = np.eye(4)
    Tc[:3, :3] = synth.rot_y(12.0)
    Tc[:3, 3] = [5.0, -3.0, 560.0]
    for dtype in (np.float32, np.float64):
        Va, Vb = V.astype(dtype), _moved(V, 25.0, [4.0, -6.0, 9.0]).astype(dtype)

        def both_frames():
            o = ctx.project(heat, K, pose[None], 0.5, frame="object", want=("t_hit", "face"))
            ctx.pose_mesh(Tc)
            c = ctx.project(heat, K, None, 0.5, frame="camera", want=("t_hit", "face"))
            return o, c, ctx.posed_vertices(dtype)

        ctx.set_mesh(Va, F).build_bvh()
        ctx.pose_mesh(Tc)
        ctx.update_vertices(Vb)
        with pytest.raises(Exception):
            ctx.project(heat, K, None, 0.5, frame="camera")         # the camera-frame copy is stale until posed again
        o1, c1, p1 = both_frames()
        ctx.set_mesh(Vb, F).build_bvh()
        o2, c2, p2 = both_frames()
        assert o1["hits"] == o2["hits"] > 0 and c1["hits"] == c2["hits"] > 0
        for a, b in ((o1, o2), (c1, c2)):
            assert np.array_equal(a["face"], b["face"])
            assert np.array_equal(a["t_hit"].view(np.uint32), b["t_hit"].view(np.uint32))
        assert np.array_equal(p1, p2)
        with pytest.raises(ValueError):
            ctx.update_vertices(Vb[:-1])
        with pytest.raises(ValueError):
            ctx.update_vertices(Vb.astype(np.float64 if dtype == np.float32 else np.float32))


def test_ray_tracing_facade_refits_a_moved_mesh(tmp_path):
    """The reference hands ray_tracing the same model at a new pose on every capture (run.py:109-110): the facade
    refits instead of rebuilding, with the outputs of a rebuild."""
    from defectproj import defect_projection as dpj
    K, H, W = synth.camera_720p()
    d = synth.write_scene_dir(str(tmp_path), K, (H, W))
    heat = synth.gaussian_heatmap((H, W), dtype=np.float64)
    V, F = synth.param_mesh(60, 40, seed=11)
    P = synth.fixed_pose()
    V1 = V.astype(np.float64) @ P[:3, :3].T + P[:3, 3]
    V2 = _moved(V, 15.0, [2.0, 1.0, -3.0]) @ P[:3, :3].T + P[:3, 3]
    dpj.ray_tracing(d, dpj.TriangleMesh(V1, F), heat, K, 0.5)
    a, ma = dpj.ray_tracing(d, dpj.TriangleMesh(V2, F), heat, K, 0.5)           # refit
    dpj._SCENE["F"] = None
    b, mb = dpj.ray_tracing(d, dpj.TriangleMesh(V2, F), heat, K, 0.5)           # rebuild
    assert len(a.points) == len(b.points) > 0
    assert np.array_equal(a.points, b.points) and np.array_equal(a.colors, b.colors)
    assert np.array_equal(a.face_ids, b.face_ids) and np.array_equal(ma.vertices, mb.vertices)
