"""GPU parity for the point-to-plane ICP row: dp_icp_point_to_plane against the numpy restatement
of Open3D's registration_icp (oracle.icp_point_to_plane; parity unpinned at the Open3D boundary).  Float64
throughout; the two differ only in the order of the sums of the 6x6 normal equations, so the bar is 1e-9 absolute
on the transform (mm / radians), exact correspondences and iteration counts."""
import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu


def _cloud_with_normals(nu=40, nv=25, seed=4):
    V, F = synth.param_mesh(nu, nv, seed=seed)
    V = V.astype(np.float64)
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    return V, vn


def _rigid(rz, rx, t):
    T = np.eye(4)
    T[:3, :3] = synth.rot_z(rz) @ synth.rot_x(rx)
    T[:3, 3] = t
    return T


@pytest.mark.parametrize("case", [dict(rz=2.0, rx=-1.5, t=[0.8, -0.5, 0.6], dist=5.0, step=3),
                                  dict(rz=0.5, rx=0.3, t=[0.1, 0.2, -0.3], dist=1.0, step=1),
                                  dict(rz=4.0, rx=2.0, t=[2.0, 1.0, -1.5], dist=8.0, step=2)])
def test_icp_equals_oracle_and_recovers_the_pose(ctx, orc, case):
    V, vn = _cloud_with_normals()
    T = _rigid(case["rz"], case["rx"], case["t"])
    src = orc.transform_points(V[::case["step"]], np.linalg.inv(T))
    src = src + np.random.default_rng(1).normal(scale=0.01, size=src.shape)
    got = ctx.icp_point_to_plane(src, V, vn, case["dist"], want_correspondence=True)
    Tr, fit, rmse, it, corr = orc.icp_point_to_plane(src, V, vn, case["dist"])
    assert got["iterations"] == it
    assert got["fitness"] == fit
    assert abs(got["inlier_rmse"] - rmse) <= 1e-9 * max(1.0, rmse)
    assert np.abs(got["transformation"] - Tr).max() <= 1e-9
    assert np.array_equal(got["correspondence"], corr)
    assert np.abs(got["transformation"] - T).max() < 0.05          # and it is the pose that was applied


def test_icp_with_initial_transform_and_single_iteration(ctx, orc):
    """predict_z_axis_adjustment's probes: max_iteration = 1 from an explicit init (:654-660)."""
    V, vn = _cloud_with_normals(30, 20, seed=6)
    T = _rigid(1.0, 0.5, [0.5, 0.2, -8.0])
    src = orc.transform_points(V[::2], np.linalg.inv(T))
    init = np.eye(4)
    init[2, 3] = -7.0
    got = ctx.icp_point_to_plane(src, V, vn, 4.0, init=init, max_iteration=1)
    Tr, fit, rmse, it, _ = orc.icp_point_to_plane(src, V, vn, 4.0, init=init, max_iteration=1)
    assert got["iterations"] == it == 1 and got["fitness"] == fit
    assert np.abs(got["transformation"] - Tr).max() <= 1e-9 and abs(got["inlier_rmse"] - rmse) <= 1e-9
    zero = ctx.icp_point_to_plane(src, V, vn, 4.0, init=init, max_iteration=0)
    assert zero["iterations"] == 0 and np.array_equal(zero["transformation"], init)


def test_icp_edge_cases(ctx, orc):
    V, vn = _cloud_with_normals(20, 12, seed=2)
    far = V[:50] + 1e4                                             # nothing within the distance: identity updates
    r = ctx.icp_point_to_plane(far, V, vn, 1.0)
    assert r["fitness"] == 0.0 and r["inlier_rmse"] == 0.0 and np.array_equal(r["transformation"], np.eye(4))
    assert r["iterations"] == 1                                    # |d fitness| = |d rmse| = 0 after one iteration
    r = ctx.icp_point_to_plane(np.zeros((0, 3)), V, vn, 1.0)
    assert r["fitness"] == 0.0 and r["iterations"] >= 0
    same = ctx.icp_point_to_plane(V, V, vn, 1.0)                   # already aligned
    assert same["fitness"] == 1.0 and same["inlier_rmse"] == 0.0
    assert np.allclose(same["transformation"], np.eye(4), atol=1e-12)
    with pytest.raises(ValueError):
        ctx.icp_point_to_plane(V, V, vn[:-1], 1.0)
    with pytest.raises(ValueError):
        ctx.icp_point_to_plane(V, V, vn, 0.0)
    # run to run reproducibility (fixed reduction order)
    T = _rigid(3.0, -2.0, [1.0, 1.0, 1.0])
    src = orc.transform_points(V, np.linalg.inv(T))
    a = ctx.icp_point_to_plane(src, V, vn, 6.0)
    b = ctx.icp_point_to_plane(src, V, vn, 6.0)
    assert np.array_equal(a["transformation"], b["transformation"]) and a["inlier_rmse"] == b["inlier_rmse"]


def test_registration_icp_facade_mirrors_open3d_call(ctx, orc):
    from defectproj import pose_estimation as pe
    from defectproj.defect_projection import PointCloud
    V, vn = _cloud_with_normals()
    T = _rigid(2.0, -1.0, [0.5, 0.5, 0.5])
    source = PointCloud(orc.transform_points(V[::2], np.linalg.inv(T)))
    target = PointCloud(V, normals=vn)
    res = pe.registration_icp(source, target, 5.0, np.eye(4), pe.TransformationEstimationPointToPlane())
    Tr, fit, rmse, it, corr = orc.icp_point_to_plane(source.points, V, vn, 5.0)
    assert np.abs(res.transformation - Tr).max() <= 1e-9 and res.fitness == fit
    assert res.correspondence_set.shape == (int((corr >= 0).sum()), 2)
    assert np.array_equal(res.correspondence_set[:, 1], corr[corr >= 0])
    one = pe.registration_icp(source, target, 5.0, np.eye(4), pe.TransformationEstimationPointToPlane(),
                              pe.ICPConvergenceCriteria(max_iteration=1))
    assert one.iterations == 1
    param = {"refine_registration": {"distance_threshold": 5.0}, "run_icp": {"fitness_threshold": 2.0, "rmse_threshold": 0.0}}
    assert np.array_equal(pe.refine_registration(source, target, np.eye(4), param).transformation, res.transformation)
    with pytest.raises(RuntimeError):
        pe.registration_icp(source, PointCloud(V), 5.0)            # no normals: Open3D raises too
    # the restart loop keeps the best result and draws in the reference's order
    best = pe.improve_result(source, target, res, param, rng=np.random.default_rng(0), max_iterations=3)
    assert best.iterations == 3 and best.fitness >= res.fitness
    assert np.array_equal(pe.get_rotation_matrix_from_xyz([0.1, 0, 0]), [[1, 0, 0], [0, np.cos(0.1), -np.sin(0.1)], [0, np.sin(0.1), np.cos(0.1)]])


def test_icp_full_size_property(ctx):
    """40k x 40k points: the recovered pose undoes the applied one (size-independent property, no oracle)."""
    V, vn = _cloud_with_normals(200, 200, seed=9)
    T = _rigid(1.5, -1.0, [0.4, -0.3, 0.5])
    Ti = np.linalg.inv(T)
    src = V @ Ti[:3, :3].T + Ti[:3, 3]
    r = ctx.icp_point_to_plane(src, V, vn, 3.0)
    assert r["fitness"] == 1.0 and r["inlier_rmse"] < 1e-6
    assert np.abs(r["transformation"] - T).max() < 1e-6


def test_icp_grid_search_equals_the_scan_bit_for_bit(ctx, monkeypatch):
    """The uniform grid over the target (targets of >= 256 points) and the tiled scan of the whole target
    (DP_ICP_GRID=0) return the same correspondences, hence the same 29 sums, transform and iteration count --
    including a radius larger than a cell's minimum (coarse grid), duplicated target points (ties to the smaller
    original index) and source points outside the target's box."""
    V, vn = _cloud_with_normals(90, 60, seed=5)
    V = np.concatenate([V, V[:500]])                                # exact duplicates: equal distances
    vn = np.concatenate([vn, vn[:500]])
    T = _rigid(2.5, -1.0, [0.7, -0.4, 0.9])
    Ti = np.linalg.inv(T)
    src = V[::3] @ Ti[:3, :3].T + Ti[:3, 3]
    src = np.concatenate([src, V[:40] + 30.0, V[:40] - 500.0])     # near and far outside the box
    for dist in (0.8, 3.0, 12.0, 400.0):                           # 400: too few cells, the library scans anyway
        monkeypatch.setenv("DP_ICP_GRID", "1")
        a = ctx.icp_point_to_plane(src, V, vn, dist, want_correspondence=True)
        monkeypatch.setenv("DP_ICP_GRID", "0")
        b = ctx.icp_point_to_plane(src, V, vn, dist, want_correspondence=True)
        assert np.array_equal(a["correspondence"], b["correspondence"]), dist
        assert a["iterations"] == b["iterations"] and a["fitness"] == b["fitness"]
        assert a["inlier_rmse"] == b["inlier_rmse"]
        assert np.array_equal(a["transformation"], b["transformation"]), dist
    assert (a["correspondence"] >= 0).any()
    # NaN points: a NaN target point is nobody's neighbour, a NaN source point has none (float64 -> int32 of a NaN is
    # INT_MIN on the device: the cell coordinate must never see it)
    V2, src2 = V.copy(), src.copy()
    V2[7] = np.nan
    src2[3] = np.nan
    monkeypatch.setenv("DP_ICP_GRID", "1")
    a = ctx.icp_point_to_plane(src2, V2, vn, 3.0, want_correspondence=True, max_iteration=2)
    monkeypatch.setenv("DP_ICP_GRID", "0")
    b = ctx.icp_point_to_plane(src2, V2, vn, 3.0, want_correspondence=True, max_iteration=2)
    assert np.array_equal(a["correspondence"], b["correspondence"]) and a["correspondence"][3] == -1
    assert not (a["correspondence"] == 7).any() and np.array_equal(a["transformation"], b["transformation"])
