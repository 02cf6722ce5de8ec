"""Deterministic synthetic inputs for the defect back-projection path.

Shared by tests/, bench.py and __graft_entry__.smoke().  Nothing here touches
the GPU or the oracle; everything is numpy and seeded.  The shapes follow
SURVEY.md section 8(d):

* ``param_mesh``      closed displaced torus, V = nu*nv, F = 2*nu*nv, millimetres
* ``camera_720p`` / ``camera_wfov``   Azure-Kinect-like pinhole intrinsics
  (resolution table of the reference: datareader.py:265-282)
* ``fixed_pose`` / ``fibonacci_poses`` / ``helix_poses``   model->camera 4x4
* ``gaussian_heatmap``   analytic restatement of the reference's
  ``generate_centered_heatmap`` (src/defect_projection.py:137-155) without cv2
* ``blob_heatmap`` / ``dense_heatmap``
* ``write_scene_dir``   writes configs/camera_{ex,in}trinsics.json in the schema
  the reference loads (src/defect_projection.py:40-92)
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

__all__ = [
    "param_mesh", "camera_720p", "camera_wfov", "K_matrix", "rot_x", "rot_y",
    "rot_z", "fixed_pose", "look_at_pose", "fibonacci_poses", "helix_poses",
    "fill_frame_pose", "gaussian_heatmap", "blob_heatmap", "dense_heatmap",
    "write_scene_dir", "MESH_CONFIGS",
]

# name -> (nu, nv): F = 2*nu*nv
MESH_CONFIGS = {
    "tiny": (12, 8),          # 192 triangles
    "small": (40, 25),        # 2 000
    "c1_30k": (150, 100),     # 30 000   (T-LESS-sized, BASELINE configs[0])
    "c2_500k": (500, 500),    # 500 000  (BASELINE configs[1])
    "ns_1m": (1000, 500),     # 1 000 000 (north_star target)
    "c4_5m": (2500, 1000),    # 5 000 000 (BASELINE configs[3])
}


def param_mesh(nu: int, nv: int, seed: int = 0, R: float = 60.0, r: float = 25.0,
               amp: float = 4.0, scale: float = 1.0):
    """Closed torus with a seeded low-frequency radial displacement.

    Returns (V float32 [nu*nv, 3] in mm, F int32 [2*nu*nv, 3]).  The grid wraps
    in both directions so the surface is watertight and every vertex is shared
    by six triangles.
    """
    rng = np.random.default_rng(seed)
    k = 4
    fu = rng.integers(1, 6, size=k)
    fv = rng.integers(1, 5, size=k)
    ph = rng.uniform(0.0, 2.0 * math.pi, size=k)
    a = rng.uniform(0.3, 1.0, size=k)
    a = a / a.sum() * amp

    u = (np.arange(nu, dtype=np.float64) / nu) * 2.0 * math.pi
    v = (np.arange(nv, dtype=np.float64) / nv) * 2.0 * math.pi
    uu, vv = np.meshgrid(u, v, indexing="ij")
    disp = np.zeros_like(uu)
    for i in range(k):
        disp += a[i] * np.sin(fu[i] * uu + fv[i] * vv + ph[i])
    rr = r + disp
    x = (R + rr * np.cos(vv)) * np.cos(uu)
    y = (R + rr * np.cos(vv)) * np.sin(uu)
    z = rr * np.sin(vv)
    V = (np.stack([x, y, z], axis=-1).reshape(-1, 3) * scale).astype(np.float32)

    iu = np.arange(nu)
    iv = np.arange(nv)
    a00 = (iu[:, None] * nv + iv[None, :]).reshape(-1)
    a10 = (((iu + 1) % nu)[:, None] * nv + iv[None, :]).reshape(-1)
    a01 = (iu[:, None] * nv + ((iv + 1) % nv)[None, :]).reshape(-1)
    a11 = (((iu + 1) % nu)[:, None] * nv + ((iv + 1) % nv)[None, :]).reshape(-1)
    F = np.empty((2 * nu * nv, 3), dtype=np.int32)
    F[0::2] = np.stack([a00, a10, a11], axis=-1)
    F[1::2] = np.stack([a00, a11, a01], axis=-1)
    return V, F


def K_matrix(fx, fy, cx, cy):
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)


def camera_720p():
    """(K, H, W) of the 720p colour camera."""
    return K_matrix(610.0, 610.0, 640.0, 360.0), 720, 1280


def camera_wfov():
    """(K, H, W) of the 1024x1024 wide-FOV depth camera."""
    return K_matrix(504.0, 504.0, 512.0, 512.0), 1024, 1024


def rot_x(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_y(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_z(deg):
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def _pose(Rm, t):
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = Rm
    T[:3, 3] = t
    return T


def fixed_pose(z: float = 600.0):
    """Model->camera pose of config C1."""
    return _pose(rot_z(30.0) @ rot_y(20.0) @ rot_x(-25.0), [0.0, 0.0, z])


def look_at_pose(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """Model->camera pose of a camera at ``eye`` (model frame) looking at ``target``.
    Camera convention: +z forward, +x right, +y down (pinhole image axes)."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    upv = np.asarray(up, dtype=np.float64)
    if abs(np.dot(fwd, upv)) > 0.999:
        upv = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, upv)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    Rcm = np.stack([right, down, fwd], axis=0)       # rows = camera axes in model frame
    return _pose(Rcm, -Rcm @ eye)


def fibonacci_poses(n: int = 64, radius: float = 600.0):
    """n cameras on a Fibonacci sphere looking at the origin (config C3)."""
    out = np.empty((n, 4, 4), dtype=np.float64)
    ga = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        zc = 1.0 - 2.0 * (i + 0.5) / n
        rad = math.sqrt(max(0.0, 1.0 - zc * zc))
        th = ga * i
        eye = radius * np.array([rad * math.cos(th), rad * math.sin(th), zc])
        out[i] = look_at_pose(eye)
    return out


def helix_poses(n: int = 1024, turns: int = 16, r0: float = 500.0, r1: float = 700.0):
    """n cameras on a helix around the object (config C5)."""
    out = np.empty((n, 4, 4), dtype=np.float64)
    for i in range(n):
        s = i / max(1, n - 1)
        th = 2.0 * math.pi * turns * s
        rad = r0 + (r1 - r0) * s
        elev = math.radians(-60.0 + 120.0 * s)
        eye = rad * np.array([math.cos(elev) * math.cos(th), math.cos(elev) * math.sin(th),
                              math.sin(elev)])
        out[i] = look_at_pose(eye)
    return out


def fill_frame_pose():
    """Pose used with ``param_mesh(..., scale=6)`` for the dense full-frame
    configs (C2/C4): the camera sits inside the tube hole region and looks along
    the ring so that nearly every pixel of the 90-degree WFOV frame lands on the
    surface.  The hit fraction is measured and reported, not assumed."""
    return look_at_pose(eye=(6 * 60.0, -30.0, 6 * 8.0), target=(6 * 30.0, 6 * 52.0, 0.0))


def gaussian_heatmap(shape, max_intensity: float = 1.0, sigma: float = 50.0,
                     dtype=np.float64):
    """Centred Gaussian, max-normalised: the analytic form of the reference's
    generate_centered_heatmap (impulse -> cv2.GaussianBlur -> /max).  cv2
    truncates its kernel, so pixel values differ slightly from cv2's in the far
    tail; the golden fixtures made from the reference record its own counts."""
    H, W = shape
    cy, cx = H // 2, W // 2
    y = np.arange(H, dtype=np.float64)[:, None] - cy
    x = np.arange(W, dtype=np.float64)[None, :] - cx
    g = np.exp(-(x * x + y * y) / (2.0 * sigma * sigma)) * max_intensity
    g = g / g.max()
    return g.astype(dtype)


def blob_heatmap(shape, seed: int = 0, nblobs: int = 8, dtype=np.float32):
    """Sum of seeded Gaussians (centres anywhere, sigma in [20, 80]), max-normalised."""
    H, W = shape
    rng = np.random.default_rng(seed)
    y = np.arange(H, dtype=np.float64)[:, None]
    x = np.arange(W, dtype=np.float64)[None, :]
    g = np.zeros((H, W), dtype=np.float64)
    for _ in range(nblobs):
        cx = rng.uniform(0, W)
        cy = rng.uniform(0, H)
        s = rng.uniform(20.0, 80.0)
        a = rng.uniform(0.4, 1.0)
        g += a * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2.0 * s * s))
    g /= g.max()
    return g.astype(dtype)


def dense_heatmap(shape, dtype=np.float32):
    """Every pixel above any threshold < 1: full-frame projection."""
    return np.ones(shape, dtype=dtype)


def write_scene_dir(path, K_color, hw_color, color_to_depth=None, K_depth=None, hw_depth=None):
    """Write ``configs/camera_extrinsics.json`` and ``configs/camera_intrinsics.json``
    in the schema the reference reads (src/defect_projection.py:40-61, 76-92)."""
    os.makedirs(os.path.join(path, "configs"), exist_ok=True)
    c2d = np.eye(4) if color_to_depth is None else np.asarray(color_to_depth, dtype=np.float64)
    d2c = np.linalg.inv(c2d)

    def ext(T):
        return {"rotation_matrix": T[:3, :3].tolist(), "translation_vector": [T[:3, 3].tolist()]}

    with open(os.path.join(path, "configs", "camera_extrinsics.json"), "w") as f:
        json.dump({"color_to_depth": ext(c2d), "depth_to_color": ext(d2c)}, f)

    def intr(K, hw):
        return {"fx": float(K[0, 0]), "fy": float(K[1, 1]), "cx": float(K[0, 2]),
                "cy": float(K[1, 2]), "width": int(hw[1]), "height": int(hw[0])}

    Kd = K_color if K_depth is None else K_depth
    hd = hw_color if hw_depth is None else hw_depth
    with open(os.path.join(path, "configs", "camera_intrinsics.json"), "w") as f:
        json.dump({"color": intr(K_color, hw_color), "depth": intr(Kd, hd)}, f)
    return path
