"""GPU parity for point-cloud normal estimation (the fallback inside align_to_surface, SURVEY.md 8f #1, and the
target preparation of the ICP row, 8f #4): dp_estimate_normals against the numpy restatement of Open3D's
estimate_normals(KDTreeSearchParamHybrid) (oracle.estimate_normals; parity unpinned at the Open3D boundary).
Neighbour sets (counts) must be equal.  The covariance is summed identically on both sides; they differ in libm's
acos/cos and in the contraction of the closed-form 3x3 eigen solver only, and that solver amplifies rounding by
about 1 / gap^2 (cross products of the rows of A - lambda I).  Bar, written where it is used: 1e-9 where the smallest
eigenvalue is separated by >= 1e-2 of the largest, 1e-6 where by >= 1e-4; nearer to a double eigenvalue the normal
is not defined."""
import numpy as np
import pytest

from defectproj import synth

pytestmark = pytest.mark.gpu


def _surface(nu, nv, seed, noise=0.02):
    V, F = synth.param_mesh(nu, nv, seed=seed)
    V = V.astype(np.float64)
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    if noise:
        V = V + np.random.default_rng(seed).normal(scale=noise, size=V.shape)
    return V, vn


def _separated(covs, rel=1e-4):
    w = np.linalg.eigvalsh(covs)
    return (w[:, 1] - w[:, 0]) >= rel * np.maximum(w[:, 2], 1e-300)


def _assert_normals_close(got, ref, covs, up_to_sign=False):
    d = np.abs(got - ref).max(1)
    if up_to_sign:                                                   # exactly symmetric neighbourhoods sit on the solver's branch points
        d = np.minimum(d, np.abs(got + ref).max(1))
    well, ok = _separated(covs, 1e-2), _separated(covs, 1e-4)
    assert well.any() and d[well].max() <= 1e-9
    assert d[ok].max() <= 1e-6
    return ok


@pytest.mark.parametrize("case", [dict(nu=60, nv=40, radius=10.0, max_nn=30),      # defect_projection.py:184-185
                                  dict(nu=90, nv=60, radius=8.0, max_nn=5),        # the cap decides (pose_estimation.py:304-305)
                                  dict(nu=50, nv=30, radius=25.0, max_nn=64),      # coarse grid, long lists
                                  dict(nu=40, nv=25, radius=400.0, max_nn=30)])    # one cell: every point is a candidate
def test_normals_equal_the_oracle(ctx, orc, case):
    P, _ = _surface(case["nu"], case["nv"], seed=3)
    got, cnt = ctx.estimate_normals(P, case["radius"], case["max_nn"], want_counts=True)
    ref, rc, covs = orc.estimate_normals(P, case["radius"], case["max_nn"])
    assert np.array_equal(cnt, rc)
    assert cnt.max() <= case["max_nn"] and cnt.min() >= 1
    ok = _assert_normals_close(got, ref, covs)
    assert ok.mean() > 0.9
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() <= 1e-12


def test_normals_reference_quirks_and_orientation(ctx, orc):
    P, vn = _surface(60, 40, seed=5)
    # align_to_surface's own parameters on a millimetre cloud (:433-435): nobody has 3 neighbours within 0.1
    got, cnt = ctx.estimate_normals(P, 0.1, 30, want_counts=True)
    assert np.all(cnt == 1) and np.all(got == [0.0, 0.0, 1.0])
    # existing normals keep their side
    free = ctx.estimate_normals(P, 10.0, 30)
    out = ctx.estimate_normals(P, 10.0, 30, normals=vn)
    assert np.all((out * vn).sum(1) >= 0.0)
    flip = (free * vn).sum(1) < 0.0
    assert flip.any() and (~flip).any()                              # the solver's own sign is not the surface's
    assert np.array_equal(out[flip], -free[flip]) and np.array_equal(out[~flip], free[~flip])
    ref, _, covs = orc.estimate_normals(P, 10.0, 30, normals=vn)
    _assert_normals_close(out, ref, covs)
    keep = ctx.estimate_normals(P, 0.1, 30, normals=vn)             # identity covariance -> (0,0,1) on the old normal's side
    assert np.array_equal(keep, np.c_[np.zeros((len(P), 2)), np.where(vn[:, 2] < 0.0, -1.0, 1.0)])
    assert np.array_equal(orc.estimate_normals(P[:50], 0.1, 30, normals=vn[:50])[0], keep[:50])


def test_normals_ties_duplicates_and_bad_input(ctx, orc):
    P, _ = _surface(40, 25, seed=7, noise=0.0)
    P = np.concatenate([P, P[:200], P[:200]])                        # triplicated points: zero distances, index order decides
    got, cnt = ctx.estimate_normals(P, 12.0, 6, want_counts=True)
    ref, rc, covs = orc.estimate_normals(P, 12.0, 6)
    assert np.array_equal(cnt, rc)
    _assert_normals_close(got, ref, covs, up_to_sign=True)
    Q = P.copy()
    Q[3] = np.nan                                                    # a NaN point has no neighbours and is nobody's neighbour
    g2, c2 = ctx.estimate_normals(Q, 12.0, 6, want_counts=True)
    assert c2[3] == 0 and np.array_equal(g2[3], [0.0, 0.0, 1.0])
    assert np.array_equal(c2, orc.estimate_normals(Q, 12.0, 6)[1])
    assert ctx.estimate_normals(np.zeros((0, 3)), 1.0, 30).shape == (0, 3)
    one = ctx.estimate_normals(np.ones((1, 3)), 1.0, 30)
    assert np.array_equal(one, [[0.0, 0.0, 1.0]])
    for bad in (dict(radius=0.0, max_nn=30), dict(radius=1.0, max_nn=0), dict(radius=1.0, max_nn=65)):
        with pytest.raises(ValueError):
            ctx.estimate_normals(P, bad["radius"], bad["max_nn"])
    Q[5, 0] = np.inf
    with pytest.raises(ValueError):
        ctx.estimate_normals(Q, 1.0, 30)


def test_normals_facade_and_align_to_surface_without_normals(ctx, orc):
    from defectproj import defect_projection as dpj
    from defectproj import pose_estimation as pe
    P, vn = _surface(60, 40, seed=9)
    pcd = dpj.estimate_normals(dpj.PointCloud(P))                    # :181-186
    assert pcd.has_normals() and np.array_equal(pcd.normals, ctx.estimate_normals(P, 10.0, 30))
    pcd2 = pe.estimate_normals(dpj.PointCloud(P), {"unused": 1})     # src/pose_estimation.py:301-306
    assert np.array_equal(pcd2.normals, ctx.estimate_normals(P, 2.0, 5))
    with pytest.raises(NotImplementedError):
        dpj.PointCloud(P).estimate_normals()
    # align_to_surface estimates the missing normals itself (:431-436) and leaves them on the cloud
    target = dpj.PointCloud(P)
    defects = np.c_[P[::7] + 0.3, np.ones(len(P[::7]))]
    off, ali = dpj.align_to_surface(defects, target, offset=0.5)
    assert target.has_normals() and np.all(target.normals == [0.0, 0.0, 1.0])   # radius 0.1 on a millimetre cloud
    o2, a2 = dpj.align_to_surface(defects, dpj.PointCloud(P, normals=target.normals), offset=0.5)
    assert np.array_equal(off, o2) and np.array_equal(ali, a2)
    assert np.array_equal(off, ali + [0.0, 0.0, 0.5])


def test_normals_full_size_property(ctx):
    """250k points of the 500x500 surface: unit normals along the surface normal (no oracle at this size)."""
    P, vn = _surface(500, 500, seed=11, noise=0.0)
    got, cnt = ctx.estimate_normals(P, 2.0, 30, want_counts=True)
    assert cnt.min() >= 3 and cnt.max() == 30
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() <= 1e-12
    assert np.abs((got * vn).sum(1)).min() > 0.99
