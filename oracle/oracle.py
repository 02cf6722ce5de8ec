"""ctypes wrapper around oracle/liboracle.so  --  TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the shipped package.
See oracle/oracle.c for what each function restates (reference file:line) and
for the "parity unpinned" statement about the Open3D/Embree boundary.

Also holds ``brute_f64_numpy``: an independent, pure-numpy float64
Moeller-Trumbore closest hit used to validate the C oracle on small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

EPS32 = float(2.0 ** -23)
# tie margins as multiples of eps32 * (max |coordinate|): see DESIGN.md "tie classification"
TAU_C = 16.0
TAU_T_C = 1024.0
GRAZE = 0.02


def build(force: bool = False) -> str:
    """Compile liboracle.so with oracle/Makefile (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, f64, f32, vp = C.c_int64, C.c_double, C.c_float, C.c_void_p
        L.orc_heatmap_to_points_f64.restype = i64
        L.orc_heatmap_to_points_f64.argtypes = [vp, i64, i64, f64, vp, vp, vp]
        L.orc_heatmap_to_points_f32.restype = i64
        L.orc_heatmap_to_points_f32.argtypes = [vp, i64, i64, f32, vp, vp, vp]
        L.orc_compute_rays.restype = None
        L.orc_compute_rays.argtypes = [vp, vp, i64, f64, f64, f64, f64, vp]
        L.orc_rays_object_frame.restype = None
        L.orc_rays_object_frame.argtypes = [vp, vp, i64, vp, vp]
        L.orc_pose_vertices.restype = None
        L.orc_pose_vertices.argtypes = [vp, i64, vp, vp]
        L.orc_tri_test_f32.restype = C.c_int
        L.orc_tri_test_f32.argtypes = [vp, vp, vp, vp, vp]
        L.orc_cast_brute_f32.restype = None
        L.orc_cast_brute_f32.argtypes = [vp, vp, i64, vp, i64, vp, vp]
        L.orc_bvh_build.restype = vp
        L.orc_bvh_build.argtypes = [vp, i64, vp, i64]
        L.orc_bvh_free.restype = None
        L.orc_bvh_free.argtypes = [vp]
        L.orc_bvh_num_nodes.restype = i64
        L.orc_bvh_num_nodes.argtypes = [vp]
        L.orc_cast_bvh_f32.restype = None
        L.orc_cast_bvh_f32.argtypes = [vp, vp, i64, vp, vp]
        L.orc_cast_f64.restype = None
        L.orc_cast_f64.argtypes = [vp, vp, i64, f64, f64, f64, C.c_int, vp, vp, vp]
        L.orc_eval_face_f64.restype = None
        L.orc_eval_face_f64.argtypes = [vp, vp, vp, vp, i64, f64, vp, vp, vp]
        L.orc_accumulate.restype = None
        L.orc_accumulate.argtypes = [vp, vp, i64, vp, vp, vp, vp]
        L.orc_project_frame_f32.restype = i64
        L.orc_project_frame_f32.argtypes = [vp, vp, i64, i64, f32, vp, vp, vp, vp, vp, vp, vp]
        L.orc_num_threads.restype = C.c_int
        L.orc_prepare_heatmap_f64.argtypes = [vp, i64, i64, i64, i64, vp]
        L.orc_prepare_heatmap_f64.restype = None
        L.orc_prepare_heatmap_f32.argtypes = [vp, i64, i64, i64, i64, vp]
        L.orc_prepare_heatmap_f32.restype = None
        L.orc_num_threads.argtypes = []
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ H1 / H2 / H3
def heatmap_to_points(heat, thr):
    """-> xs int64[N], ys int64[N], I (heat dtype)[N]; restates src/defect_projection.py:165-179."""
    heat = np.ascontiguousarray(heat)
    H, W = heat.shape
    if heat.dtype == np.float32:
        fn, dt = lib().orc_heatmap_to_points_f32, np.float32
    else:
        heat = _c(heat, np.float64)
        fn, dt = lib().orc_heatmap_to_points_f64, np.float64
    n = fn(_p(heat), H, W, float(thr), None, None, None)
    xs = np.empty(n, np.int64)
    ys = np.empty(n, np.int64)
    I = np.empty(n, dt)
    fn(_p(heat), H, W, float(thr), _p(xs), _p(ys), _p(I))
    return xs, ys, I


def compute_rays(xs, ys, K):
    """-> float64 [N,3] unit rays in the camera frame; restates :196-223."""
    xs = _c(xs, np.int64)
    ys = _c(ys, np.int64)
    K = np.asarray(K, dtype=np.float64)
    out = np.empty((len(xs), 3), np.float64)
    lib().orc_compute_rays(_p(xs), _p(ys), len(xs), K[0, 0], K[1, 1], K[0, 2], K[1, 2], _p(out))
    return out


def frame_xform(K, pose):
    """16 doubles: fx fy cx cy | Rinv row-major | tinv  for a model->camera pose."""
    K = np.asarray(K, np.float64)
    pose = np.asarray(pose, np.float64)
    Rm, t = pose[:3, :3], pose[:3, 3]
    Rinv = Rm.T
    tinv = -(Rinv @ t)
    return np.concatenate([[K[0, 0], K[1, 1], K[0, 2], K[1, 2]], Rinv.reshape(-1), tinv]).astype(np.float64)


def rays_object_frame(xs, ys, xf):
    xs = _c(xs, np.int64)
    ys = _c(ys, np.int64)
    xf = _c(xf, np.float64)
    out = np.empty((len(xs), 6), np.float32)
    lib().orc_rays_object_frame(_p(xs), _p(ys), len(xs), _p(xf), _p(out))
    return out


def rays6_camera(rays_f64):
    """The float32 [N,6] tensor the reference builds at :247-251 (origin 0)."""
    r = np.zeros((len(rays_f64), 6), np.float32)
    r[:, 3:] = rays_f64.astype(np.float32)
    return r


def pose_vertices(V, T):
    V = _c(V, np.float64)
    T = _c(T, np.float64)
    out = np.empty((len(V), 3), np.float32)
    lib().orc_pose_vertices(_p(V), len(V), _p(T), _p(out))
    return out


# ------------------------------------------------------------------ H4
class Bvh:
    def __init__(self, V, F):
        self.V = _c(V, np.float32)
        self.F = _c(F, np.int32)
        self.h = lib().orc_bvh_build(_p(self.V), len(self.V), _p(self.F), len(self.F))
        self.scale = float(np.abs(self.V).max()) if len(self.V) else 1.0

    def __del__(self):
        try:
            if self.h:
                lib().orc_bvh_free(self.h)
                self.h = None
        except Exception:
            pass

    def cast_f32(self, rays6):
        rays6 = _c(rays6, np.float32)
        n = len(rays6)
        t = np.empty(n, np.float32)
        f = np.empty(n, np.int32)
        lib().orc_cast_bvh_f32(self.h, _p(rays6), n, _p(t), _p(f))
        return t, f

    def margins(self, rays6=None):
        s = self.scale
        if rays6 is not None and len(rays6):
            s = max(s, float(np.abs(np.asarray(rays6)[:, :3]).max()))
        return TAU_C * EPS32 * s, TAU_T_C * EPS32 * s

    def cast_f64(self, rays6, brute=False, tau=None, tau_t=None, graze=GRAZE):
        rays6 = _c(rays6, np.float32)
        n = len(rays6)
        a, b = self.margins(rays6)
        tau = a if tau is None else tau
        tau_t = b if tau_t is None else tau_t
        t = np.empty(n, np.float64)
        f = np.empty(n, np.int32)
        tie = np.empty(n, np.uint8)
        lib().orc_cast_f64(self.h, _p(rays6), n, tau, tau_t, graze, int(brute), _p(t), _p(f), _p(tie))
        return t, f, tie

    def eval_face(self, rays6, face, tau=None):
        rays6 = _c(rays6, np.float32)
        face = _c(face, np.int32)
        n = len(rays6)
        tau = self.margins(rays6)[0] if tau is None else tau
        t = np.empty(n, np.float64)
        m = np.empty(n, np.float64)
        c = np.empty(n, np.float64)
        lib().orc_eval_face_f64(_p(self.V), _p(self.F), _p(rays6), _p(face), n, tau, _p(t), _p(m), _p(c))
        return t, m, c

    def project_frame(self, heat, thr, K, want_acc=True):
        """Whole frame on the CPU (all threads): the timed CPU baseline."""
        heat = _c(heat, np.float32)
        H, W = heat.shape
        K = np.asarray(K, np.float64)
        K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64)
        t = np.empty(H * W, np.float32)
        f = np.empty(H * W, np.int32)
        hist = np.zeros(len(self.F), np.int32) if want_acc else None
        fmax = np.zeros(len(self.F), np.float32) if want_acc else None
        vmax = np.zeros(len(self.V), np.float32) if want_acc else None
        nh = C.c_int64(0)
        n = lib().orc_project_frame_f32(self.h, _p(heat), H, W, float(thr), _p(K4), _p(t), _p(f),
                                        _p(hist), _p(fmax), _p(vmax), C.byref(nh))
        return dict(n=n, t=t[:n], face=f[:n], hist=hist, fmax=fmax, vmax=vmax, nhits=nh.value)


def cast_brute_f32(V, F, rays6):
    V = _c(V, np.float32)
    F = _c(F, np.int32)
    rays6 = _c(rays6, np.float32)
    n = len(rays6)
    t = np.empty(n, np.float32)
    f = np.empty(n, np.int32)
    lib().orc_cast_brute_f32(_p(V), _p(F), len(F), _p(rays6), n, _p(t), _p(f))
    return t, f


def tri_test_f32(ray6, v0, v1, v2):
    ray6 = _c(ray6, np.float32)
    v0, v1, v2 = (_c(v, np.float32) for v in (v0, v1, v2))
    t = C.c_float(0)
    hit = lib().orc_tri_test_f32(_p(ray6), _p(v0), _p(v1), _p(v2), C.byref(t))
    return bool(hit), (float(t.value) if hit else float("inf"))


def accumulate(face, I, F, nV):
    face = _c(face, np.int32)
    I = _c(I, np.float32)
    F = _c(F, np.int32)
    hist = np.zeros(len(F), np.int32)
    fmax = np.zeros(len(F), np.float32)
    vmax = np.zeros(nV, np.float32)
    lib().orc_accumulate(_p(face), _p(I), len(face), _p(F), _p(hist), _p(fmax), _p(vmax))
    return hist, fmax, vmax


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n=None):
    """OpenMP threads of the following calls; default = the CPUs this process may run on (a launcher's
    OMP_NUM_THREADS=1 would otherwise make the CPU arm single-threaded).  Returns the count in effect."""
    if n is None:
        n = len(os.sched_getaffinity(0))
    lib().orc_set_num_threads(int(n))
    return num_threads()


# ------------------------------------------------------------------ independent numpy check
def brute_f64_numpy(V, F, rays6, chunk=256):
    """Pure-numpy float64 Moeller-Trumbore closest hit (min t, ties -> smaller face id).
    O(N*F) memory-chunked; for validating the C oracle on small cases only."""
    V = np.asarray(V, np.float32).astype(np.float64)
    F = np.asarray(F, np.int64)
    r = np.asarray(rays6, np.float32).astype(np.float64)
    v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    e1, e2 = v1 - v0, v2 - v0
    n = len(r)
    tout = np.full(n, np.inf)
    fout = np.full(n, -1, np.int32)
    for s in range(0, n, chunk):
        o = r[s:s + chunk, None, :3]
        d = r[s:s + chunk, None, 3:]
        P = np.cross(d, e2[None])
        det = np.einsum("fk,nfk->nf", e1, P)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            T = o - v0[None]
            u = np.einsum("nfk,nfk->nf", T, P) * inv
            Q = np.cross(T, e1[None])
            v = np.einsum("nfk,nfk->nf", np.broadcast_to(d, Q.shape), Q) * inv
            t = np.einsum("fk,nfk->nf", e2, Q) * inv
        ok = (det != 0) & (u >= 0) & (v >= 0) & (1.0 - u - v >= 0) & (t >= 0)
        t = np.where(ok, t, np.inf)
        j = np.argmin(t, axis=1)          # first minimum => smaller face id on ties
        tt = t[np.arange(len(j)), j]
        tout[s:s + chunk] = tt
        fout[s:s + chunk] = np.where(np.isfinite(tt), j, -1)
    return tout, fout


# ------------------------------------------------------------------ depth-image projection path (numpy, float64)
def heatmap_to_point3d(heat, depth, K, thr=0.1):
    """Vectorised restatement of src/defect_projection.py:359-395 (same selection, order and arithmetic)."""
    heat = np.asarray(heat)
    depth = np.asarray(depth)
    K = np.asarray(K, np.float64)
    H, W = heat.shape
    Hd, Wd = depth.shape
    h = min(H, Hd)
    w = min(W, Wd)
    with np.errstate(divide="ignore", invalid="ignore"):
        if heat.dtype == np.float32:
            # :384 on a float32 map is a float32 division; its rounded quotient is compared with the (float64) threshold
            # the way numpy 1.26.4 compares an np.float32 scalar with a Python float -- in float64 -- and stored widened
            maxv = np.max(heat) if heat.size else np.float32(1.0)
            inten = (heat[:h, :w] / maxv).astype(np.float64)
        else:
            maxv = np.max(heat).astype(np.float64) if heat.size else np.float64(1.0)
            inten = heat[:h, :w].astype(np.float64) / maxv
    d = depth[:h, :w]
    ys, xs = np.nonzero((inten > thr) & (d > 0))
    dd = d[ys, xs].astype(np.float64)
    x3 = (xs - K[0, 2]) * dd / K[0, 0]
    y3 = (ys - K[1, 2]) * dd / K[1, 1]
    return np.stack([x3, y3, dd * 0.98, inten[ys, xs]], axis=1)


def nearest_points(query, target):
    """Exact float64 nearest neighbour, first minimum (smaller index) on ties: stands in for KDTreeFlann (k = 1)."""
    q = np.asarray(query, np.float64)[:, :3]
    t = np.asarray(target, np.float64)
    idx = np.empty(len(q), np.int32)
    for s in range(0, len(q), 512):
        d = q[s:s + 512, None, :] - t[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        idx[s:s + 512] = np.argmin(d2, axis=1)
    return idx


def align_to_surface(query, target, normals, offset):
    idx = nearest_points(query, target)
    t = np.asarray(target, np.float64)
    return t[idx] + np.asarray(normals, np.float64)[idx] * offset, t[idx], idx


# ------------------------------------------------------------------ before / after the path
def prepare_heatmap(data, H, W):
    """DataReader.get_heatmap (datareader.py:639-675): min-max normalise, cv2 INTER_LINEAR resize to min(H, W)^2,
    centre in a zero H x W float64 frame.  The resize is the C restatement pinned against the real cv2."""
    data = np.asarray(data)
    dt = np.float32 if data.dtype == np.float32 else np.float64
    d = np.ascontiguousarray(data, dtype=dt)
    out = np.empty((H, W), np.float64)
    fn = lib().orc_prepare_heatmap_f32 if dt == np.float32 else lib().orc_prepare_heatmap_f64
    fn(_p(d), d.shape[0], d.shape[1], H, W, _p(out))
    return out


_JET_SEG = {
    0: ((0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)),
    1: ((0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)),
    2: ((0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)),
}


def jet_lut(N=256):
    """matplotlib.colors._create_lookup_table for the 'jet' segment data (public algorithm; matplotlib is absent)."""
    lut = np.empty((N, 3), np.float64)
    for c, data in _JET_SEG.items():
        a = np.array(data, dtype=np.float64)
        x, y0, y1 = a[:, 0] * (N - 1), a[:, 1], a[:, 2]
        xind = np.linspace(0, N - 1, N)
        ind = np.searchsorted(x, xind)[1:-1]
        dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        lut[:, c] = np.clip(np.concatenate([[y1[0]], dist * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]]), 0.0, 1.0)
    return lut


def jet(x):
    """RGB of plt.get_cmap('jet')(x)[:, :3] (Colormap.__call__: x*N, truncation, x == 1 -> N-1, NaN -> (0,0,0))."""
    x = np.array(x, dtype=np.float64)
    bad = np.isnan(x)
    with np.errstate(invalid="ignore"):
        xa = x * 256
        xa[xa < 0] = -1
        xa[xa == 256] = 255
        xi = np.clip(np.where(bad, 0, xa), -1, 256).astype(int)
    xi[xi > 255] = 255          # "over" colour = lut[N-1]
    xi[xi < 0] = 0              # "under" colour = lut[0]
    rgb = jet_lut()[xi]
    rgb[bad] = 0.0
    return rgb


def transform_points(p, T):
    """Open3D PointCloud.transform: (T @ [p, 1])[:3] / w, evaluated left to right in float64."""
    p = np.asarray(p, np.float64).reshape(-1, 3)
    T = np.asarray(T, np.float64)
    h = [((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3] for r in range(4)]
    return np.stack([h[0] / h[3], h[1] / h[3], h[2] / h[3]], axis=1)


def pack_hits(intensity, face=None, point64=None, T=None):
    """Selection of the hits (:259-264), colours of create_intersection_pcd (:286-291), optional transform (run.py:118)."""
    I = np.asarray(intensity).astype(np.float64)
    sel = np.ones(len(I), bool) if face is None else np.asarray(face) >= 0
    Is = I[sel]
    out = {"index": np.nonzero(sel)[0], "intensity": Is}
    if len(Is):
        with np.errstate(invalid="ignore", divide="ignore"):
            out["colors"] = jet((Is - np.min(Is)) / (np.max(Is) - np.min(Is)))
    else:
        out["colors"] = np.zeros((0, 3))
    if point64 is not None:
        pts = np.asarray(point64, np.float64).reshape(-1, 3)[sel]
        out["points"] = transform_points(pts, T) if T is not None else pts
    return out


# ------------------------------------------------------------------ point-to-plane ICP
def _nn_within(q, t, max_dist):
    idx = np.empty(len(q), np.int64)
    d2 = np.empty(len(q), np.float64)
    for s in range(0, len(q), 256):
        d = q[s:s + 256, None, :] - t[None, :, :]
        dd = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        j = np.argmin(dd, axis=1)
        idx[s:s + 256] = j
        d2[s:s + 256] = dd[np.arange(len(j)), j]
    ok = d2 <= max_dist * max_dist
    return idx, d2, ok


def icp_point_to_plane(source, target, normals, max_dist, init=None, max_iteration=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """Open3D's RegistrationICP + TransformationEstimationPointToPlane restated in numpy (PARITY UNPINNED: open3d
    is absent; algorithm as published in Registration.cpp / TransformationEstimation.cpp / Eigen.cpp).
    Returns (T, fitness, inlier_rmse, iterations, correspondence)."""
    src = np.asarray(source, np.float64).reshape(-1, 3)
    tp = np.asarray(target, np.float64).reshape(-1, 3)
    tn = np.asarray(normals, np.float64).reshape(-1, 3)
    T = np.eye(4) if init is None else np.array(init, np.float64)
    pcd = src if np.array_equal(T, np.eye(4)) else transform_points(src, T)

    def evaluate(p):
        idx, d2, ok = _nn_within(p, tp, max_dist)
        k = int(ok.sum())
        return idx, ok, (k / len(p) if len(p) else 0.0), (np.sqrt(d2[ok].sum() / k) if k else 0.0)

    idx, ok, fit, rmse = evaluate(pcd)
    it = 0
    while it < max_iteration:
        update = np.eye(4)
        if ok.any():
            s, t, n = pcd[ok], tp[idx[ok]], tn[idx[ok]]
            r = np.einsum("ij,ij->i", s - t, n)
            J = np.concatenate([np.cross(s, n), n], axis=1)
            JTJ, JTr = J.T @ J, J.T @ r
            det = np.linalg.det(JTJ)
            if np.isfinite(det) and abs(det) >= 1e-6:
                x = np.linalg.solve(JTJ, -JTr)
                ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
                Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
                Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
                Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
                update[:3, :3] = Rz @ Ry @ Rx
                update[:3, 3] = x[3:]
        T = update @ T
        pcd = transform_points(pcd, update)
        pf, pr = fit, rmse
        idx, ok, fit, rmse = evaluate(pcd)
        it += 1
        if abs(pf - fit) < rel_fitness and abs(pr - rmse) < rel_rmse:
            break
    return T, fit, rmse, it, np.where(ok, idx, -1).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# Normal estimation (PointCloud.estimate_normals with KDTreeSearchParamHybrid): numpy / pure-Python restatement
# of Open3D 0.18's published algorithm as recalled (EstimateNormals.cpp, KDTreeFlann::SearchHybrid,
# utility::ComputeCovariance, FastEigen3x3 after Eberly's robust 3x3 eigensolver).  PARITY UNPINNED: Open3D is
# absent offline.  Call sites: /root/reference/src/defect_projection.py:181-186, :431-436,
# /root/reference/src/pose_estimation.py:301-306.
def _eigenvector0(A, ev):
    import math
    r0 = (A[0][0] - ev, A[0][1], A[0][2])
    r1 = (A[0][1], A[1][1] - ev, A[1][2])
    r2 = (A[0][2], A[1][2], A[2][2] - ev)

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    cs = [cross(r0, r1), cross(r0, r2), cross(r1, r2)]
    ds = [c[0] * c[0] + c[1] * c[1] + c[2] * c[2] for c in cs]
    imax, dmax = 0, ds[0]
    if ds[1] > dmax:
        dmax, imax = ds[1], 1
    if ds[2] > dmax:
        imax = 2
    s = math.sqrt(ds[imax])
    if s == 0.0:                                    # 0 / 0 in IEEE arithmetic (Python would raise)
        return (float("nan"),) * 3
    return tuple(x / s for x in cs[imax])


def _eigenvector1(A, e0, ev):
    import math
    if abs(e0[0]) > abs(e0[1]):
        inv = 1.0 / math.sqrt(e0[0] * e0[0] + e0[2] * e0[2])
        U = (-e0[2] * inv, 0.0, e0[0] * inv)
    else:
        inv = 1.0 / math.sqrt(e0[1] * e0[1] + e0[2] * e0[2])
        U = (0.0, e0[2] * inv, -e0[1] * inv)
    V = (e0[1] * U[2] - e0[2] * U[1], e0[2] * U[0] - e0[0] * U[2], e0[0] * U[1] - e0[1] * U[0])
    AU = tuple(A[r][0] * U[0] + A[r][1] * U[1] + A[r][2] * U[2] for r in range(3))
    AV = tuple(A[r][0] * V[0] + A[r][1] * V[1] + A[r][2] * V[2] for r in range(3))
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]   # noqa: E731
    m00, m01, m11 = dot(U, AU) - ev, dot(U, AV), dot(V, AV) - ev
    a00, a01, a11 = abs(m00), abs(m01), abs(m11)
    if a00 >= a11:
        if max(a00, a01) > 0.0:
            if a00 >= a01:
                m01 /= m00
                m00 = 1.0 / math.sqrt(1.0 + m01 * m01)
                m01 *= m00
            else:
                m00 /= m01
                m01 = 1.0 / math.sqrt(1.0 + m00 * m00)
                m00 *= m01
            return tuple(m01 * U[k] - m00 * V[k] for k in range(3))
        return U
    if max(a11, a01) > 0.0:
        if a11 >= a01:
            m01 /= m11
            m11 = 1.0 / math.sqrt(1.0 + m01 * m01)
            m01 *= m11
        else:
            m11 /= m01
            m01 = 1.0 / math.sqrt(1.0 + m11 * m11)
            m11 *= m01
        return tuple(m11 * U[k] - m01 * V[k] for k in range(3))
    return U


def fast_eigen3x3(C):
    """Eigenvector of the smallest eigenvalue of the symmetric 3x3 matrix C (nested lists / array)."""
    import math
    C = [[float(C[r][c]) for c in range(3)] for r in range(3)]
    mx = max(max(row) for row in C)
    if mx == 0.0:
        return (0.0, 0.0, 0.0)
    A = [[C[r][c] / mx for c in range(3)] for r in range(3)]
    norm = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2]
    cross = lambda a, b: (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])   # noqa: E731
    if norm > 0.0:
        q = (A[0][0] + A[1][1] + A[2][2]) / 3.0
        b00, b11, b22 = A[0][0] - q, A[1][1] - q, A[2][2] - q
        p = math.sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0)
        c00 = b11 * b22 - A[1][2] * A[1][2]
        c01 = A[0][1] * b22 - A[1][2] * A[0][2]
        c02 = A[0][1] * A[1][2] - b11 * A[0][2]
        det = (b00 * c00 - A[0][1] * c01 + A[0][2] * c02) / (p * p * p)
        half_det = min(max(det * 0.5, -1.0), 1.0)
        angle = math.acos(half_det) / 3.0
        two_thirds_pi = 2.09439510239319549
        beta2 = math.cos(angle) * 2.0
        beta0 = math.cos(angle + two_thirds_pi) * 2.0
        beta1 = -(beta0 + beta2)
        e0, e1, e2 = q + p * beta0, q + p * beta1, q + p * beta2
        if half_det >= 0.0:
            v2 = _eigenvector0(A, e2)
            if e2 < e0 and e2 < e1:
                return v2
            v1 = _eigenvector1(A, v2, e1)
            if e1 < e0 and e1 < e2:
                return v1
            return cross(v1, v2)
        v0 = _eigenvector0(A, e0)
        if e0 < e1 and e0 < e2:
            return v0
        v1 = _eigenvector1(A, v0, e1)
        if e1 < e0 and e1 < e2:
            return v1
        return cross(v0, v1)
    if C[0][0] < C[1][1] and C[0][0] < C[2][2]:
        return (1.0, 0.0, 0.0)
    if C[1][1] < C[0][0] and C[1][1] < C[2][2]:
        return (0.0, 1.0, 0.0)
    return (0.0, 0.0, 1.0)


def hybrid_neighbours(points, i, radius, max_nn):
    """Indices of the max_nn nearest points of points[i] (itself included) with squared distance < radius^2,
    ordered by (distance, index); every operation of the distance individually rounded (float64)."""
    d = points[i] - points
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    cand = np.nonzero(d2 < radius * radius)[0]
    order = np.lexsort((cand, d2[cand]))
    return cand[order][:max_nn]


def neighbourhood_covariance(points, idx):
    """utility::ComputeCovariance: nine cumulants summed in the order of idx, then E[xx^T] - E[x]E[x]^T."""
    c = [0.0] * 9
    for j in idx:
        x, y, z = float(points[j, 0]), float(points[j, 1]), float(points[j, 2])
        c[0] += x; c[1] += y; c[2] += z                                   # noqa: E702
        c[3] += x * x; c[4] += x * y; c[5] += x * z                       # noqa: E702
        c[6] += y * y; c[7] += y * z; c[8] += z * z                       # noqa: E702
    m = float(len(idx))
    c = [v / m for v in c]
    return [[c[3] - c[0] * c[0], c[4] - c[0] * c[1], c[5] - c[0] * c[2]],
            [c[4] - c[0] * c[1], c[6] - c[1] * c[1], c[7] - c[1] * c[2]],
            [c[5] - c[0] * c[2], c[7] - c[1] * c[2], c[8] - c[2] * c[2]]]


def estimate_normals(points, radius, max_nn=30, normals=None):
    """(normals [n,3], neighbour counts [n], covariances [n,3,3]) for small clouds (O(n^2) distances)."""
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    n = len(points)
    has = normals is not None and len(normals) == n and n > 0
    out = np.zeros((n, 3))
    counts = np.zeros(n, np.int32)
    covs = np.zeros((n, 3, 3))
    for i in range(n):
        idx = hybrid_neighbours(points, i, radius, max_nn) if np.all(np.isfinite(points[i])) else np.zeros(0, np.int64)
        counts[i] = len(idx)
        C = neighbourhood_covariance(points, idx) if len(idx) >= 3 else [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
        covs[i] = C
        v = fast_eigen3x3(C)
        if v[0] * v[0] + v[1] * v[1] + v[2] * v[2] == 0.0:
            v = tuple(normals[i]) if has else (0.0, 0.0, 1.0)
        if has and v[0] * normals[i][0] + v[1] * normals[i][1] + v[2] * normals[i][2] < 0.0:
            v = (-v[0], -v[1], -v[2])
        out[i] = v
    return out, counts, covs
