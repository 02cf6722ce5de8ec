"""Minimal stand-ins for `open3d` and `matplotlib`, used ONLY by make_golden.py to run the
reference's src/defect_projection.py unmodified in a container that has neither package.

Scope: exactly the attributes that module touches on the ray_tracing path
(src/defect_projection.py:1-10 imports, :225-317, :527-563).  RaycastingScene.cast_rays is
answered by the CPU oracle's float32 closest hit (oracle semantic A); everything else is a
plain container or a numpy restatement of documented Open3D / matplotlib behaviour:
  * legacy TriangleMesh.transform: v' = (T @ [v,1])[:3] / w in float64
  * t.geometry.TriangleMesh.from_legacy: vertices -> float32, triangles -> int64
  * core.Tensor(ndarray, dtype=Float32): cast to float32
  * cast_rays: dict with 't_hit' float32 [N], inf on miss
  * matplotlib LinearSegmentedColormap('jet', N=256) lookup incl. the NaN -> "bad" colour
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


# --------------------------------------------------------------------------- matplotlib.cm jet
_JET = {
    "red": ((0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)),
    "green": ((0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)),
    "blue": ((0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)),
}


def _create_lookup_table(N, data):
    adata = np.array(data, dtype=np.float64)
    x, y0, y1 = adata[:, 0], adata[:, 1], adata[:, 2]
    x = x * (N - 1)
    xind = np.linspace(0, N - 1, N)
    ind = np.searchsorted(x, xind)[1:-1]
    distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.concatenate([[y1[0]], distance * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    return np.clip(lut, 0.0, 1.0)


class _Jet:
    N = 256

    def __init__(self):
        lut = np.ones((self.N + 3, 4), dtype=np.float64)
        for i, c in enumerate(("red", "green", "blue")):
            lut[:self.N, i] = _create_lookup_table(self.N, _JET[c])
        lut[self.N] = lut[0]                 # under
        lut[self.N + 1] = lut[self.N - 1]    # over
        lut[self.N + 2] = (0.0, 0.0, 0.0, 0.0)   # bad (NaN)
        self._lut = lut

    def __call__(self, X):
        xa = np.array(X, dtype=np.float64, copy=True)
        bad = np.isnan(xa)
        with np.errstate(invalid="ignore"):
            xa *= self.N
            xa[xa < 0] = -1
            xa[xa == self.N] = self.N - 1
            np.clip(xa, -1, self.N, out=xa)
            xi = xa.astype(int)
        xi[xi > self.N - 1] = self.N + 1
        xi[xi < 0] = self.N
        xi[bad] = self.N + 2
        return self._lut[xi]


def _make_matplotlib():
    mpl = types.ModuleType("matplotlib")
    cm = types.ModuleType("matplotlib.cm")
    plt = types.ModuleType("matplotlib.pyplot")
    jet = _Jet()

    def get_cmap(name):
        assert name == "jet"
        return jet

    cm.get_cmap = get_cmap
    plt.get_cmap = get_cmap
    mpl.cm = cm
    mpl.pyplot = plt
    return mpl, cm, plt


# --------------------------------------------------------------------------- open3d
class _Vec:
    """Vector3dVector / Vector2iVector: np.asarray() gives the array back."""

    def __init__(self, a, dtype):
        self._a = np.array(a, dtype=dtype)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __len__(self):
        return len(self._a)


class TriangleMesh:
    def __init__(self, vertices=None, triangles=None):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.array(vertices, dtype=np.float64)
        self.triangles = np.zeros((0, 3), np.int32) if triangles is None else np.array(triangles, dtype=np.int32)
        self._tn = False
        self._vn = False

    def has_triangle_normals(self):
        return self._tn

    def has_vertex_normals(self):
        return self._vn

    def compute_triangle_normals(self):
        self._tn = True
        return self

    def compute_vertex_normals(self):
        self._vn = True
        self._tn = True
        return self

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        v = self.vertices
        h = np.concatenate([v, np.ones((len(v), 1))], axis=1) @ T.T
        self.vertices = h[:, :3] / h[:, 3:4]
        return self

    def paint_uniform_color(self, c):
        return self


class PointCloud:
    def __init__(self):
        self.points = _Vec(np.zeros((0, 3)), np.float64)
        self.colors = _Vec(np.zeros((0, 3)), np.float64)
        self.normals = _Vec(np.zeros((0, 3)), np.float64)

    def has_normals(self):
        return len(self.normals) > 0


class KDTreeFlann:
    """search_knn_vector_3d(p, 1) answered by an exact float64 scan (first minimum = smaller index)."""

    def __init__(self, pcd):
        self._p = np.asarray(pcd.points, dtype=np.float64)

    def search_knn_vector_3d(self, query, knn):
        assert knn == 1
        d = self._p - np.asarray(query, dtype=np.float64)[None, :]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        i = int(np.argmin(d2))
        return 1, [i], [float(d2[i])]


class LineSet:
    def __init__(self):
        self.points = _Vec(np.zeros((0, 3)), np.float64)
        self.lines = _Vec(np.zeros((0, 2)), np.int32)
        self.colors = _Vec(np.zeros((0, 3)), np.float64)

    def paint_uniform_color(self, c):
        self.colors = _Vec(np.tile(np.asarray(c, np.float64), (len(self.lines), 1)), np.float64)
        return self


class PinholeCameraIntrinsic:
    def __init__(self, width=0, height=0, fx=0.0, fy=0.0, cx=0.0, cy=0.0):
        self.width, self.height = width, height
        self.intrinsic_matrix = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)

    def set_intrinsics(self, width, height, fx, fy, cx, cy):
        self.__init__(width, height, fx, fy, cx, cy)


class _TTensor:
    def __init__(self, a, dtype=None):
        self._a = np.asarray(a, dtype=np.float32 if dtype == "Float32" else None)

    def numpy(self):
        return self._a


class _TMesh:
    @staticmethod
    def from_legacy(mesh):
        m = _TMesh()
        m.v = np.asarray(mesh.vertices, dtype=np.float64).astype(np.float32)
        m.f = np.asarray(mesh.triangles).astype(np.int64)
        return m


class RaycastingScene:
    def __init__(self):
        self._bvh = None

    def add_triangles(self, tmesh):
        from oracle import oracle as orc
        self._bvh = orc.Bvh(tmesh.v, tmesh.f.astype(np.int32))
        return 0

    def cast_rays(self, rays):
        r = np.ascontiguousarray(rays.numpy(), dtype=np.float32)
        t, f = self._bvh.cast_f32(r)
        return {"t_hit": _TTensor(t), "primitive_ids": _TTensor(f.astype(np.uint32))}


def _make_open3d():
    o3d = types.ModuleType("open3d")
    geometry = types.ModuleType("open3d.geometry")
    geometry.TriangleMesh = TriangleMesh
    geometry.PointCloud = PointCloud
    geometry.LineSet = LineSet
    geometry.KDTreeFlann = KDTreeFlann
    geometry.KDTreeSearchParamHybrid = lambda radius=0.0, max_nn=0: None
    utility = types.ModuleType("open3d.utility")
    utility.Vector3dVector = lambda a: _Vec(a, np.float64)
    utility.Vector2iVector = lambda a: _Vec(a, np.int32)
    camera = types.ModuleType("open3d.camera")
    camera.PinholeCameraIntrinsic = PinholeCameraIntrinsic
    core = types.ModuleType("open3d.core")
    core.Tensor = _TTensor
    core.Dtype = types.SimpleNamespace(Float32="Float32")
    t = types.ModuleType("open3d.t")
    tgeo = types.ModuleType("open3d.t.geometry")
    tgeo.TriangleMesh = _TMesh
    tgeo.RaycastingScene = RaycastingScene
    t.geometry = tgeo
    o3d.geometry, o3d.utility, o3d.camera, o3d.core, o3d.t = geometry, utility, camera, core, t
    return {"open3d": o3d, "open3d.geometry": geometry, "open3d.utility": utility,
            "open3d.camera": camera, "open3d.core": core, "open3d.t": t, "open3d.t.geometry": tgeo}


def install():
    mods = _make_open3d()
    mpl, cm, plt = _make_matplotlib()
    mods.update({"matplotlib": mpl, "matplotlib.cm": cm, "matplotlib.pyplot": plt})
    for k, v in mods.items():
        sys.modules.setdefault(k, v)
