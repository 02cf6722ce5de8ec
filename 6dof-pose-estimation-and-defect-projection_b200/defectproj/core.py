"""Thin object wrapper over the C ABI (include/defectproj.h).

Host (numpy) arguments are staged by the library; CUDA tensors (torch) are passed by
address and stay on the device.  Nothing here computes: it marshals pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (DP_DEVICE, DP_E_ARG, DP_E_NOMEM, DP_E_STATE, DP_F32, DP_F64, DP_FRAME_CAMERA,
                   DP_FRAME_OBJECT, DP_HOST, RaysOut, Stats)

__all__ = ["Context", "DefectProjError", "FRAME_OBJECT", "FRAME_CAMERA"]

FRAME_OBJECT, FRAME_CAMERA = DP_FRAME_OBJECT, DP_FRAME_CAMERA
_PINNED_KEEPALIVE = []


class DefectProjError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[defectproj {code}] {msg}")
        self.code = code


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _frame(frame):
    if frame in ("object", FRAME_OBJECT):
        return FRAME_OBJECT
    if frame in ("camera", FRAME_CAMERA):
        return FRAME_CAMERA
    raise ValueError(f"frame must be 'object' or 'camera', got {frame!r}")


class DevView:
    """__cuda_array_interface__ view of library-owned device memory (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Context:
    """One dp_ctx: a mesh, its BVH(s), accumulators and scratch on one B200.  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.dp_create(int(device), C.byref(h))
        if rc != 0:
            msg = self._L.dp_last_error(None).decode()
            raise ValueError(msg) if rc == DP_E_ARG else DefectProjError(rc, msg)
        self._h = h
        self.device = int(device)
        self.nV = self.nF = 0

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.dp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc == 0:
            return
        msg = self._L.dp_last_error(self._h).decode()
        if rc == DP_E_ARG:
            raise ValueError(msg)
        if rc == DP_E_NOMEM:
            raise MemoryError(msg)
        raise DefectProjError(rc, msg)

    @staticmethod
    def _stream(stream):
        if stream is None:
            return None
        if hasattr(stream, "cuda_stream"):
            return C.c_void_p(stream.cuda_stream)
        return C.c_void_p(int(stream))

    def synchronize(self, stream=None):
        self._check(self._L.dp_synchronize(self._h, self._stream(stream)))

    # ------------------------------------------------------------------ mesh / BVH
    def set_mesh(self, V, F, stream=None):
        """V [nV,3] float32 or float64, F [nF,3] int32 (numpy or CUDA tensors)."""
        if _is_torch(V):
            import torch
            if not (V.is_cuda and F.is_cuda):
                raise ValueError("tensor meshes must live on the device; pass numpy arrays for host data")
            V = V.contiguous()
            F = F.to(torch.int32).contiguous()
            vd = DP_F64 if V.dtype == torch.float64 else DP_F32
            if V.dtype not in (torch.float32, torch.float64):
                V = V.float()
            nV, nF, mem = V.shape[0], F.shape[0], DP_DEVICE
        else:
            V = np.asarray(V)
            vd = DP_F64 if V.dtype == np.float64 else DP_F32
            V = np.ascontiguousarray(V, dtype=np.float64 if vd == DP_F64 else np.float32)
            F = np.ascontiguousarray(F, dtype=np.int32)
            if V.ndim != 2 or (V.size and V.shape[1] != 3) or F.ndim != 2 or (F.size and F.shape[1] != 3):
                raise ValueError("V must be [nV,3] and F [nF,3]")
            nV, nF, mem = len(V), len(F), DP_HOST
        self._check(self._L.dp_set_mesh(self._h, _ptr(V), vd, nV, _ptr(F), nF, mem, self._stream(stream)))
        self.nV, self.nF = nV, nF
        return self

    def build_bvh(self, stream=None):
        self._check(self._L.dp_build_bvh(self._h, self._stream(stream)))
        return self

    def update_vertices(self, V, stream=None):
        """New positions for the vertices of set_mesh (same count, dtype and faces): refit instead of a rebuild."""
        if _is_torch(V):
            import torch
            if not V.is_cuda or V.dtype not in (torch.float32, torch.float64):
                raise ValueError("vertex tensors must be float32/float64 CUDA tensors")
            V = V.contiguous()
            vd, n, mem = (DP_F64 if V.dtype == torch.float64 else DP_F32), V.shape[0], DP_DEVICE
        else:
            V = np.ascontiguousarray(V)
            if V.dtype not in (np.float32, np.float64):
                V = V.astype(np.float64)
            V = V.reshape(-1, 3)
            vd, n, mem = (DP_F64 if V.dtype == np.float64 else DP_F32), len(V), DP_HOST
        self._check(self._L.dp_update_vertices(self._h, _ptr(V), vd, n, mem, self._stream(stream)))
        return self

    def pose_mesh(self, T, stream=None):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.dp_pose_mesh(self._h, _ptr(T), self._stream(stream)))
        return self

    def posed_vertices(self, dtype=np.float32, stream=None):
        dtype = np.dtype(dtype)
        out = np.empty((self.nV, 3), dtype)
        self._check(self._L.dp_get_posed_vertices(self._h, _ptr(out), DP_F64 if dtype == np.float64 else DP_F32,
                                                  DP_HOST, self._stream(stream)))
        return out

    def posed_vertices_device(self, dtype=np.float32, stream=None):
        """The posed vertices as a CUDA tensor [nV,3] owned by the caller (a device-to-device copy: nothing crosses PCIe)."""
        import torch
        dtype = np.dtype(dtype)
        tdt = torch.float64 if dtype == np.float64 else torch.float32
        out = torch.empty((self.nV, 3), dtype=tdt, device=f"cuda:{self.device}")
        if stream is None:
            stream = torch.cuda.current_stream(out.device)
        self._check(self._L.dp_get_posed_vertices(self._h, _ptr(out), DP_F64 if dtype == np.float64 else DP_F32, DP_DEVICE,
                                                  self._stream(stream)))
        return out

    @staticmethod
    def pinned_array(shape, dtype):
        """A page-locked numpy array (host staging for inputs that arrive in pageable memory)."""
        import torch
        t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
        a = t.numpy()
        _PINNED_KEEPALIVE.append(t)
        return a

    # ------------------------------------------------------------------ H1
    def compact(self, heat, thr, want_intensity=True, stream=None):
        """np.where(heat > thr) on the GPU.  heat [H,W] or [B,H,W], float32/float64 numpy.
        Returns (pixel uint32 [N] = frame*H*W + y*W + x, intensity float32 [N] | None, counts int64 [B])."""
        heat = np.asarray(heat)
        if heat.dtype not in (np.float32, np.float64):
            heat = heat.astype(np.float64)
        heat = np.ascontiguousarray(heat)
        if heat.ndim == 2:
            heat = heat[None]
        if heat.ndim != 3:
            raise ValueError("heat must be [H,W] or [B,H,W]")
        B, H, W = heat.shape
        cap = heat.size
        pix = np.empty(cap, np.uint32)
        inten = np.empty(cap, np.float32) if want_intensity else None
        counts = np.zeros(B, np.int64)
        n = C.c_int64(0)
        self._check(self._L.dp_compact(self._h, _ptr(heat), DP_F64 if heat.dtype == np.float64 else DP_F32, B, H, W,
                                       float(thr), _ptr(pix), _ptr(inten), cap, C.byref(n), _ptr(counts), DP_HOST,
                                       self._stream(stream)))
        k = n.value
        return pix[:k], (inten[:k] if want_intensity else None), counts

    # ------------------------------------------------------------------ H2
    def compute_rays(self, xs, ys, K, stream=None):
        xs = np.ascontiguousarray(xs, dtype=np.int32)
        ys = np.ascontiguousarray(ys, dtype=np.int32)
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        out = np.empty((len(xs), 3), np.float64)
        self._check(self._L.dp_compute_rays(self._h, _ptr(xs), _ptr(ys), len(xs), _ptr(K), _ptr(out), DP_HOST,
                                            self._stream(stream)))
        return out

    @staticmethod
    def frame_xform(K, pose=None):
        """The 16 per-frame constants the kernels use (fx fy cx cy | Rinv | tinv)."""
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        P = None if pose is None else np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        out = np.empty(16, np.float64)
        _lib.load().dp_frame_xform(_ptr(K), _ptr(P), _ptr(out))
        return out

    # ------------------------------------------------------------------ H4
    def cast_rays(self, rays6, frame="object", want_face=True, stream=None):
        """Closest hit of explicit float32 rays [N,6] -> (t_hit float32 [N], face int32 [N])."""
        fr = _frame(frame)
        if _is_torch(rays6):
            import torch
            r = rays6.contiguous().float()
            n = r.shape[0]
            t = torch.empty(n, dtype=torch.float32, device=r.device)
            f = torch.empty(n, dtype=torch.int32, device=r.device) if want_face else None
            self._check(self._L.dp_cast_rays(self._h, fr, _ptr(r), n, _ptr(t), _ptr(f), DP_DEVICE,
                                             self._stream(stream if stream is not None else torch.cuda.current_stream())))
            return t, f
        r = np.ascontiguousarray(rays6, dtype=np.float32)
        if r.ndim != 2 or (r.size and r.shape[1] != 6):
            raise ValueError("rays6 must be [N,6]")
        n = len(r)
        t = np.empty(n, np.float32)
        f = np.empty(n, np.int32) if want_face else None
        self._check(self._L.dp_cast_rays(self._h, fr, _ptr(r), n, _ptr(t), _ptr(f), DP_HOST, self._stream(stream)))
        return t, f

    # ------------------------------------------------------------------ fused path
    def project(self, heat, K, poses=None, thr=0.5, frame="object", accumulate=True,
                want=("pixel", "intensity", "t_hit", "face", "point"), cap=None, out=None, stream=None):
        """Host-buffer end-to-end call: heat (numpy [H,W] or [B,H,W]) in, per-ray arrays out.

        `out` may map output names to preallocated (ideally pinned) numpy arrays to be filled in place;
        otherwise arrays for the names in `want` are allocated.  Returns
        dict(n=rays, hits=..., pixel=..., intensity=..., t_hit=..., face=..., point=..., point64=...)
        with only the requested arrays, each cut to n.
        """
        fr = _frame(frame)
        heat = np.asarray(heat)
        if heat.dtype not in (np.float32, np.float64):
            heat = heat.astype(np.float64)
        heat = np.ascontiguousarray(heat)
        if heat.ndim == 2:
            heat = heat[None]
        B, H, W = heat.shape
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, 9)
        P = None
        if fr == FRAME_OBJECT:
            if poses is None:
                raise ValueError("object-frame projection needs the model->camera pose(s)")
            P = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 16)
            if len(P) != B:
                raise ValueError(f"{B} frames but {len(P)} poses")
        dtypes = {"pixel": np.uint32, "intensity": np.float32, "t_hit": np.float32, "face": np.int32,
                  "point": np.float32, "point64": np.float64}
        ro = RaysOut()
        if out is not None:
            arrs = dict(out)
            for k, a in arrs.items():
                if a.dtype != dtypes[k] or not a.flags.c_contiguous:
                    raise ValueError(f"out[{k!r}] must be C-contiguous {np.dtype(dtypes[k]).name}")
            cap = min(len(a) for a in arrs.values()) if arrs else 0
        else:
            cap = heat.size if cap is None else int(cap)
            arrs = {k: np.empty((cap, 3) if k.startswith("point") else (cap,), dtypes[k]) for k in want}
        ro.cap = cap
        for k, a in arrs.items():
            setattr(ro, k, _ptr(a))
        n, nh = C.c_int64(0), C.c_int64(0)
        self._check(self._L.dp_project(self._h, fr, _ptr(heat), DP_F64 if heat.dtype == np.float64 else DP_F32, B, H, W,
                                       float(thr), _ptr(K), len(K), _ptr(P), int(bool(accumulate)), C.byref(ro),
                                       C.byref(n), C.byref(nh), DP_HOST, self._stream(stream)))
        res = {k: v[:n.value] for k, v in arrs.items()}
        res["n"], res["hits"] = n.value, nh.value
        return res

    def project_device(self, heat, K, poses=None, thr=0.5, frame="object", accumulate=True, out=None,
                       sync=False, stream=None):
        """Device-resident call: `heat` is a CUDA tensor [B,H,W] (float32/float64); `out` maps
        names to preallocated CUDA tensors.  Asynchronous on the current torch stream unless sync."""
        import torch
        fr = _frame(frame)
        if not heat.is_cuda:
            raise ValueError("project_device needs a CUDA tensor")
        heat = heat.contiguous()
        if heat.dim() == 2:
            heat = heat[None]
        B, H, W = heat.shape
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, 9)
        P = None
        if fr == FRAME_OBJECT:
            P = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 16)
            if len(P) != B:
                raise ValueError(f"{B} frames but {len(P)} poses")
        ro = RaysOut()
        ro.cap = 0
        if out:
            caps = []
            for k, tsr in out.items():
                setattr(ro, k, _ptr(tsr))
                if k != "counts":
                    caps.append(tsr.shape[0])
            ro.cap = min(caps) if caps else 0
        n, nh = C.c_int64(0), C.c_int64(0)
        st = self._stream(stream if stream is not None else torch.cuda.current_stream())
        self._check(self._L.dp_project(self._h, fr, _ptr(heat), DP_F64 if heat.dtype == torch.float64 else DP_F32, B, H,
                                       W, float(thr), _ptr(K), len(K), _ptr(P), int(bool(accumulate)),
                                       C.byref(ro) if out else None, C.byref(n) if sync else None,
                                       C.byref(nh) if sync else None, DP_DEVICE, st))
        return (n.value, nh.value) if sync else None

    # ------------------------------------------------------------------ depth-image projection path
    @staticmethod
    def _depth_u16(depth):
        depth = np.asarray(depth)
        if depth.dtype != np.uint16:
            raise ValueError("depth image must be uint16 (cv2.IMREAD_UNCHANGED of a Kinect depth PNG)")
        if depth.ndim != 2:
            raise ValueError("depth image must be 2-D")
        return np.ascontiguousarray(depth)

    def depth_backproject(self, heat, depth, K, thr=0.1, stream=None):
        """heatmap_to_point3d on the GPU -> float64 [n,4] rows (x3d, y3d, z3d, intensity), row-major pixel order."""
        heat = np.asarray(heat)
        if heat.dtype not in (np.float32, np.float64):
            heat = heat.astype(np.float64)
        heat = np.ascontiguousarray(heat)
        if heat.ndim != 2:
            raise ValueError("heatmap must be 2-D")
        depth = self._depth_u16(depth)
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        cap = heat.size
        out = np.empty((cap, 4), np.float64)
        n = C.c_int64(0)
        self._check(self._L.dp_depth_backproject(self._h, _ptr(heat), DP_F64 if heat.dtype == np.float64 else DP_F32,
                                                 heat.shape[0], heat.shape[1], _ptr(depth), depth.shape[0], depth.shape[1],
                                                 _ptr(K), float(thr), _ptr(out), cap, C.byref(n), DP_HOST, self._stream(stream)))
        return out[:n.value]

    def calc_coordinates(self, xs, ys, depth, K, stream=None):
        xs = np.ascontiguousarray(xs, dtype=np.int32)
        ys = np.ascontiguousarray(ys, dtype=np.int32)
        depth = self._depth_u16(depth)
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        out = np.empty((len(xs), 3), np.float64)
        valid = np.empty(len(xs), np.uint8)
        self._check(self._L.dp_calc_coordinates(self._h, _ptr(xs), _ptr(ys), len(xs), _ptr(depth), depth.shape[0], depth.shape[1],
                                                _ptr(K), _ptr(out), _ptr(valid), DP_HOST, self._stream(stream)))
        return out, valid.astype(bool)

    def align_to_surface(self, query, target, normals=None, offset=0.1, stream=None):
        """Exact nearest target point of every query point (first three columns), float64.
        Returns (offset_points | None, aligned_points, idx)."""
        q = np.ascontiguousarray(query, dtype=np.float64)
        if q.ndim != 2 or q.shape[1] < 3:
            raise ValueError("query must be [n, >=3]")
        t = np.ascontiguousarray(target, dtype=np.float64).reshape(-1, 3)
        nr = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
        if nr is not None and len(nr) != len(t):
            raise ValueError("normals and target points differ in length")
        n = len(q)
        offs = np.empty((n, 3), np.float64) if nr is not None else None
        ali = np.empty((n, 3), np.float64)
        idx = np.empty(n, np.int32)
        self._check(self._L.dp_align_to_surface(self._h, _ptr(q), q.shape[1], n, _ptr(t), _ptr(nr), len(t), float(offset),
                                                _ptr(offs), _ptr(ali), _ptr(idx), DP_HOST, self._stream(stream)))
        return offs, ali, idx

    # ------------------------------------------------------------------ before / after the path (8f #3, #2)
    def prepare_heatmap(self, data, H, W, out_dtype=np.float64, stream=None):
        """DataReader.get_heatmap's array work (datareader.py:658-674) on the GPU: min-max normalise, cv2-exact
        INTER_LINEAR resize to min(H, W)^2, centred in a zero [H, W] frame.  `data` may be a CUDA tensor, in which
        case a CUDA tensor is returned (ready for project_device)."""
        if _is_torch(data):
            import torch
            if data.dtype not in (torch.float32, torch.float64) or data.dim() != 2 or not data.is_contiguous():
                raise ValueError("device heatmap data must be a contiguous 2-D float32/float64 tensor")
            odt = torch.float64 if np.dtype(out_dtype) == np.float64 else torch.float32
            out = torch.empty((int(H), int(W)), dtype=odt, device=data.device)
            if stream is None:
                stream = torch.cuda.current_stream(data.device)
            self._check(self._L.dp_prepare_heatmap(self._h, _ptr(data), DP_F64 if data.dtype == torch.float64 else DP_F32,
                                                   data.shape[0], data.shape[1], int(H), int(W), _ptr(out),
                                                   DP_F64 if odt == torch.float64 else DP_F32, DP_DEVICE, self._stream(stream)))
            return out
        data = np.asarray(data)
        if data.ndim != 2 or data.size == 0:
            raise ValueError("heatmap data must be a non-empty 2-D array")
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        data = np.ascontiguousarray(data)
        out_dtype = np.dtype(out_dtype)
        if out_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("out_dtype must be float32 or float64")
        out = np.empty((int(H), int(W)), out_dtype)
        self._check(self._L.dp_prepare_heatmap(self._h, _ptr(data), DP_F64 if data.dtype == np.float64 else DP_F32,
                                               data.shape[0], data.shape[1], int(H), int(W), _ptr(out),
                                               DP_F64 if out_dtype == np.float64 else DP_F32, DP_HOST, self._stream(stream)))
        return out

    def transform_points(self, points, T, stream=None):
        """PointCloud.transform on the GPU (float64); returns a new [n,3] array (a CUDA tensor is transformed in place)."""
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        if _is_torch(points):
            import torch
            if points.dtype != torch.float64 or not points.is_contiguous():
                raise ValueError("device points must be a contiguous float64 tensor")
            if stream is None:
                stream = torch.cuda.current_stream(points.device)
            self._check(self._L.dp_transform_points(self._h, _ptr(points), points.numel() // 3, _ptr(T), DP_DEVICE,
                                                    self._stream(stream)))
            return points
        p = np.array(points, dtype=np.float64, order="C").reshape(-1, 3)
        self._check(self._L.dp_transform_points(self._h, _ptr(p), len(p), _ptr(T), DP_HOST, self._stream(stream)))
        return p

    def jet_lut(self):
        lut = np.empty((256, 3), np.float64)
        self._L.dp_jet_lut(_ptr(lut))
        return lut

    def pack_hits(self, intensity, face=None, pixel=None, point64=None, T=None,
                  want=("points", "colors", "face", "pixel", "intensity"), stream=None):
        """The viewer payload of one projection: the rays that hit (face >= 0; all when face is None) in ray order,
        their jet colours (create_intersection_pcd, :286-291) and their points moved by T (run.py:118)."""
        I = np.asarray(intensity)
        if I.dtype not in (np.float32, np.float64):
            I = I.astype(np.float64)
        I = np.ascontiguousarray(I).reshape(-1)
        n = len(I)
        face = None if face is None else np.ascontiguousarray(face, dtype=np.int32).reshape(-1)
        pixel = None if pixel is None else np.ascontiguousarray(pixel, dtype=np.uint32).reshape(-1)
        p64 = None if point64 is None else np.ascontiguousarray(point64, dtype=np.float64).reshape(-1, 3)
        for a in (face, pixel, p64):
            if a is not None and len(a) != n:
                raise ValueError("per-ray arrays differ in length")
        Tm = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        cap = n
        out = {}
        if "points" in want and p64 is not None:
            out["points"] = np.empty((cap, 3), np.float64)
        if "colors" in want:
            out["colors"] = np.empty((cap, 3), np.float64)
        if "face" in want and face is not None:
            out["face"] = np.empty(cap, np.int32)
        if "pixel" in want:
            out["pixel"] = np.empty(cap, np.uint32)
        if "intensity" in want:
            out["intensity"] = np.empty(cap, np.float64)
        m = C.c_int64(0)
        self._check(self._L.dp_pack_hits(self._h, _ptr(I), DP_F64 if I.dtype == np.float64 else DP_F32, _ptr(face), _ptr(pixel),
                                         _ptr(p64), n, _ptr(Tm), _ptr(out.get("points")), _ptr(out.get("colors")),
                                         _ptr(out.get("face")), _ptr(out.get("pixel")), _ptr(out.get("intensity")), cap,
                                         C.byref(m), DP_HOST, self._stream(stream)))
        out = {k: v[:m.value] for k, v in out.items()}
        out["m"] = m.value
        return out

    def pack_hits_device(self, intensity, face=None, pixel=None, point64=None, T=None,
                         want=("points", "colors", "face", "pixel", "intensity"), stream=None):
        """pack_hits on CUDA tensors (the per-ray outputs of project_device): nothing but the count crosses PCIe.
        intensity: float32/float64 [n]; face: int32 [n]; pixel: int32 [n] (bits of the uint32 index); point64:
        float64 [n,3].  Returns CUDA tensors cut to the number of selected rays, plus 'm'."""
        import torch
        if not intensity.is_cuda or intensity.dtype not in (torch.float32, torch.float64):
            raise ValueError("intensity must be a float32/float64 CUDA tensor")
        dev, n = intensity.device, intensity.numel()
        chk = {"face": (face, torch.int32, n), "pixel": (pixel, torch.int32, n), "point64": (point64, torch.float64, 3 * n)}
        for name, (t, dt, cnt) in chk.items():
            if t is not None and (not t.is_cuda or t.dtype != dt or t.numel() != cnt or not t.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous {dt} CUDA tensor matching the intensities")
        if not intensity.is_contiguous():
            intensity = intensity.contiguous()
        Tm = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        out = {}
        if "points" in want and point64 is not None:
            out["points"] = torch.empty((n, 3), dtype=torch.float64, device=dev)
        if "colors" in want:
            out["colors"] = torch.empty((n, 3), dtype=torch.float64, device=dev)
        if "face" in want and face is not None:
            out["face"] = torch.empty(n, dtype=torch.int32, device=dev)
        if "pixel" in want:
            out["pixel"] = torch.empty(n, dtype=torch.int32, device=dev)
        if "intensity" in want:
            out["intensity"] = torch.empty(n, dtype=torch.float64, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev)
        m = C.c_int64(0)
        self._check(self._L.dp_pack_hits(self._h, _ptr(intensity), DP_F64 if intensity.dtype == torch.float64 else DP_F32,
                                         _ptr(face), _ptr(pixel), _ptr(point64), n, _ptr(Tm), _ptr(out.get("points")),
                                         _ptr(out.get("colors")), _ptr(out.get("face")), _ptr(out.get("pixel")),
                                         _ptr(out.get("intensity")), n, C.byref(m), DP_DEVICE, self._stream(stream)))
        out = {k: v[:m.value] for k, v in out.items()}
        out["m"] = m.value
        return out

    # ------------------------------------------------------------------ upstream of the path (8f #4)
    def icp_point_to_plane(self, source, target, target_normals, max_correspondence_distance, init=None,
                           max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6, want_correspondence=False,
                           stream=None):
        """registration_icp with TransformationEstimationPointToPlane on the GPU.
        Returns dict(transformation [4,4], fitness, inlier_rmse, iterations[, correspondence [n] int32])."""
        src = np.ascontiguousarray(source, dtype=np.float64).reshape(-1, 3)
        tp = np.ascontiguousarray(target, dtype=np.float64).reshape(-1, 3)
        tn = np.ascontiguousarray(target_normals, dtype=np.float64).reshape(-1, 3)
        if len(tn) != len(tp):
            raise ValueError("target normals and points differ in length")
        T0 = None if init is None else np.ascontiguousarray(init, dtype=np.float64).reshape(16)
        T = np.empty(16, np.float64)
        fit, rmse, it = C.c_double(0), C.c_double(0), C.c_int(0)
        corr = np.empty(len(src), np.int32) if want_correspondence else None
        self._check(self._L.dp_icp_point_to_plane(self._h, _ptr(src), len(src), _ptr(tp), _ptr(tn), len(tp),
                                                  float(max_correspondence_distance), _ptr(T0), int(max_iteration),
                                                  float(relative_fitness), float(relative_rmse), _ptr(T), C.byref(fit),
                                                  C.byref(rmse), C.byref(it), _ptr(corr), DP_HOST, self._stream(stream)))
        out = dict(transformation=T.reshape(4, 4), fitness=fit.value, inlier_rmse=rmse.value, iterations=it.value)
        if want_correspondence:
            out["correspondence"] = corr
        return out

    def estimate_normals(self, points, radius, max_nn=30, normals=None, want_counts=False, stream=None):
        """PointCloud.estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) on the GPU.  `normals`: the cloud's
        existing normals (the new ones keep their side) or None.  Returns normals [n,3] (and the neighbour counts)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        has = normals is not None and len(normals) == len(pts) and len(pts) > 0
        out = np.array(normals, dtype=np.float64).reshape(-1, 3) if has else np.empty((len(pts), 3), np.float64)
        cnt = np.empty(len(pts), np.int32) if want_counts else None
        self._check(self._L.dp_estimate_normals(self._h, _ptr(pts), len(pts), float(radius), int(max_nn), _ptr(out),
                                                int(has), _ptr(cnt), DP_HOST, self._stream(stream)))
        return (out, cnt) if want_counts else out

    # ------------------------------------------------------------------ H6 / H7
    def accum_reset(self, stream=None):
        self._check(self._L.dp_accum_reset(self._h, self._stream(stream)))

    def accum_get(self, stream=None):
        hist = np.empty(self.nF, np.int32)
        fmax = np.empty(self.nF, np.float32)
        vmax = np.empty(self.nV, np.float32)
        self._check(self._L.dp_accum_get(self._h, _ptr(hist), _ptr(fmax), _ptr(vmax), DP_HOST, self._stream(stream)))
        return hist, fmax, vmax

    def accum_flush(self, stream=None):
        """Bring the per-vertex maxima up to date on the device (they are derived from the per-face maxima)."""
        self._check(self._L.dp_accum_flush(self._h, self._stream(stream)))

    def accum_device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.dp_accum_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def accum_layout(self):
        """(base address, hist offset, fmax offset, vmax offset, total bytes) of the one accumulator allocation."""
        base = C.c_void_p()
        a, b, c, n = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.dp_accum_layout(self._h, C.byref(base), C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return base.value, a.value, b.value, c.value, n.value

    # ------------------------------------------------------------------ multi-GPU helpers
    def set_ray_shard(self, rank: int = 0, world: int = 1):
        """Trace only block `rank` of `world` of every frame's compacted ray list (single-frame ray sharding)."""
        self._check(self._L.dp_set_ray_shard(self._h, int(rank), int(world)))
        return self

    @staticmethod
    def shard_slots(rank: int, world: int, n_rays: int, H: int, W: int, nframes: int = 1):
        """Output slots [lo, hi) of shard `rank` of `world` for a projection that selected n_rays pixels."""
        lo, hi = C.c_int64(0), C.c_int64(0)
        rc = _lib.load().dp_shard_slots(int(rank), int(world), int(n_rays), int(nframes), int(H), int(W), C.byref(lo), C.byref(hi))
        if rc != 0:
            raise ValueError("shard_slots: bad arguments")
        return lo.value, hi.value

    def pack_records_device(self, t_hit, face, pixel=None, point=None, first=0, n=None, out=None, count_async=None,
                            sync=True, stream=None):
        """Hit records of the rays [first, first+n) of project_device's outputs: rows (pixel, t_hit bits, face[, x, y, z
        bits]) of the rays that hit, in ray order, as an int32 CUDA tensor [m, 3|6].  `out` may be a preallocated
        [>= n, 3|6] int32 tensor; `count_async`: pinned int64 tensor [1] that receives m on the stream (sync=False)."""
        import torch
        if n is None:
            n = face.numel() - first
        w = 6 if point is not None else 3
        if out is None:
            out = torch.empty((max(int(n), 1), w), dtype=torch.int32, device=face.device)
        if out.shape[1] != w or out.dtype != torch.int32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous int32 [cap, {w}] tensor")
        m = C.c_int64(0)
        if stream is None:
            stream = torch.cuda.current_stream(face.device)
        self._check(self._L.dp_pack_records(self._h, _ptr(pixel), _ptr(t_hit), _ptr(face), _ptr(point), int(n), int(first),
                                            _ptr(out), out.shape[0], C.byref(m) if sync else None,
                                            _ptr(count_async) if count_async is not None else None, DP_DEVICE,
                                            self._stream(stream)))
        return out[:m.value] if sync else out

    # ------------------------------------------------------------------ exchange over peer-mapped memory (csrc/peer.cu)
    def peer_export(self, record_bytes: int = 0, result_rays: int = 0):
        """Allocate this context's exchange window; returns its 64-byte CUDA IPC handle (bytes)."""
        buf = (C.c_ubyte * _lib.DP_PEER_HANDLE_BYTES)()
        n = C.c_int64(0)
        self._check(self._L.dp_peer_export(self._h, int(record_bytes), int(result_rays), buf, C.byref(n)))
        self.peer_window_bytes = n.value
        return bytes(buf)

    def peer_open(self, rank: int, world: int, handles):
        """handles: the handles of all ranks in rank order (bytes of world * 64, or a list of bytes)."""
        blob = b"".join(handles) if isinstance(handles, (list, tuple)) else bytes(handles)
        if len(blob) != world * _lib.DP_PEER_HANDLE_BYTES:
            raise ValueError("peer_open needs one 64-byte handle per rank")
        self._check(self._L.dp_peer_open(self._h, int(rank), int(world), C.c_char_p(blob)))
        self.peer_rank, self.peer_world = int(rank), int(world)
        return self

    def peer_open_local(self, rank: int, contexts):
        """The same for contexts of this process (a list in rank order; contexts[rank] is self)."""
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        self._check(self._L.dp_peer_open_local(self._h, int(rank), len(contexts), arr))
        self.peer_rank, self.peer_world = int(rank), len(contexts)
        return self

    def peer_close(self):
        self._check(self._L.dp_peer_close(self._h))

    def peer_region(self, what: int, slot: int):
        """(device address, bytes) of a region of the own window (DP_PEER_* of include/defectproj.h)."""
        ptr, n = C.c_void_p(), C.c_int64(0)
        self._check(self._L.dp_peer_window(self._h, int(what), int(slot), C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def peer_tensor(self, what: int, slot: int, row_words: int = 3):
        """Torch view (no copy) of a region of the own window: snapshot int32 [words]; records int32 [rows, row_words];
        record count int64 [1]; t_hit float32 [cap]; face int32 [cap]; point float32 [cap, 3]."""
        import torch
        ptr, n = self.peer_region(what, slot)
        dev = f"cuda:{self.device}"
        if n == 0:
            return None
        if what == _lib.DP_PEER_REC_COUNT:
            return torch.as_tensor(DevView(ptr, 1, "<i8"), device=dev)
        if what in (_lib.DP_PEER_T_HIT, _lib.DP_PEER_POINT):
            t = torch.as_tensor(DevView(ptr, n // 4, "<f4"), device=dev)
            return t.view(-1, 3) if what == _lib.DP_PEER_POINT else t
        t = torch.as_tensor(DevView(ptr, n // 4, "<i4"), device=dev)
        if what == _lib.DP_PEER_RECORDS:
            rows = (n // 4) // row_words
            return t[:rows * row_words].view(rows, row_words)
        return t

    def peer_snapshot(self, slot: int, reset: bool = True, stream=None):
        self._check(self._L.dp_peer_snapshot(self._h, int(slot), int(bool(reset)), self._stream(stream)))

    def peer_combine(self, slot: int, total=None, gather_root: int = -1, gathered=None, count_async=None, stream=None):
        """ONE launch: barrier over the ranks, totals (+)= / max= every rank's snapshot `slot`, and on rank `gather_root`
        the record slots of all ranks land in `gathered` (int32 CUDA tensor [cap_rows, row_words]) in rank order;
        `count_async` (pinned or CUDA int64 [1]) receives the total row count."""
        rows, words = (gathered.shape[0], gathered.shape[1]) if gathered is not None else (0, 3)
        self._check(self._L.dp_peer_combine(self._h, int(slot), _ptr(total), int(gather_root), _ptr(gathered), int(rows), int(words),
                                            _ptr(count_async) if count_async is not None else None, self._stream(stream)))

    def peer_results(self, slot: int = -1, with_points: bool = False):
        self._check(self._L.dp_peer_results(self._h, int(slot), int(bool(with_points))))

    def peer_status(self):
        e = C.c_int(0)
        self._check(self._L.dp_peer_status(self._h, C.byref(e)))
        return e.value

    # ------------------------------------------------------------------ instrumentation
    def set_stats(self, on: bool):
        self._check(self._L.dp_set_stats(self._h, int(bool(on))))

    def stats(self):
        st = Stats()
        self._check(self._L.dp_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_ if k != "reserved"}

    def set_timing(self, on: bool):
        """Stage events inside dp_project on/off (off: no last_timings, a few microseconds less per call)."""
        self._check(self._L.dp_set_timing(self._h, int(bool(on))))

    def last_timings(self):
        ms = (C.c_float * 4)()
        self._check(self._L.dp_last_timings(self._h, ms))
        return {"compact_ms": ms[0], "raygen_ms": ms[1], "trace_ms": ms[2], "total_ms": ms[3]}

    def dump_bvh(self, frame="object"):
        fr = _frame(frame)
        nn, nt = C.c_int64(0), C.c_int64(0)
        self._check(self._L.dp_debug_dump_bvh(self._h, fr, None, C.byref(nn), None, C.byref(nt)))
        nodes = np.empty((nn.value, 20), np.uint32)
        tris = np.empty((nt.value, 12), np.float32)
        self._check(self._L.dp_debug_dump_bvh(self._h, fr, _ptr(nodes), C.byref(nn), _ptr(tris), C.byref(nt)))
        return nodes, tris

    def ray_node_counts(self, n):
        out = np.empty(int(n), np.uint32)
        self._check(self._L.dp_debug_ray_nodes(self._h, _ptr(out), int(n)))
        return out

    def radix_sort(self, keys, vals):
        keys = np.array(keys, dtype=np.uint32, copy=True)
        vals = np.array(vals, dtype=np.uint32, copy=True)
        self._check(self._L.dp_debug_radix_sort(self._h, _ptr(keys), _ptr(vals), len(keys)))
        return keys, vals

    def morton_codes(self):
        codes = np.empty(self.nF, np.uint32)
        self._check(self._L.dp_debug_morton(self._h, _ptr(codes)))
        return codes
