"""Multi-frame defect accumulation: the loop around the back-projection in /root/reference/run.py
(:100-120 first detection, :175-206 later detections).

Per detection the reference
  1. poses the mesh with ``inv(current_transformation)``            (transform_object, :109-110 / :179-181)
  2. calls ``ray_tracing`` (threshold 0.75)                          (:113-116 / :187-193)
  3. moves every EARLIER hit cloud by ``relative_transformation = inv(current) @ previous``   (:183-184, :196-197)
  4. moves the new cloud into the depth camera's frame with ``color_to_depth``                  (:118 / :200)
  5. ships all clouds to the viewer                                                             (:205)

``DefectTracker`` keeps that behaviour (the clouds and their re-posing run on the GPU through
``PointCloud.transform`` -> ``dp_transform_points``) and, because the BVH lives in the OBJECT frame here, also
keeps what the reference cannot: a persistent per-face hit histogram / max intensity over all detections, which
needs no re-posing at all (a face id is the same in every frame).
"""
from __future__ import annotations

import numpy as np

from . import defect_projection as _dp

__all__ = ["DefectTracker", "relative_transformation"]


def relative_transformation(current_transformation, previous_transformation):
    """run.py:183-184."""
    return np.linalg.inv(current_transformation) @ previous_transformation


class DefectTracker:
    def __init__(self, target_mesh, color_intrinsics, color_to_depth, heatmap_threshold=0.75):
        self.V, self.F = _dp._mesh_arrays(target_mesh)
        self.K = _dp._K(color_intrinsics)
        self.color_to_depth = np.asarray(color_to_depth, dtype=np.float64)
        self.threshold = heatmap_threshold
        self.intersection_pcds = []
        self.previous_transformation = None
        self.mesh_in_camera = None            # the posed model in the DEPTH camera's frame (what update_dash_data gets)
        # the tracker's own context: its device accumulators are never reset, so the per-face histogram and the
        # per-face / per-vertex maxima persist across detections without leaving the GPU
        from .core import Context
        self.ctx = Context(_dp.get_context().device)
        self.ctx.set_mesh(self.V, self.F)
        self.ctx.build_bvh()
        self.ctx.accum_reset()

    @property
    def hist(self):
        return self.ctx.accum_get()[0]

    @property
    def fmax(self):
        return self.ctx.accum_get()[1]

    @property
    def vmax(self):
        return self.ctx.accum_get()[2]

    def add_detection(self, heatmap, current_transformation):
        """One defect-detection frame.  ``current_transformation`` is the ICP result (camera -> model); the mesh is
        posed by its inverse and by inv(color_to_depth), exactly the product ray_tracing applies (:549-550)."""
        cur = np.asarray(current_transformation, dtype=np.float64)
        ctx = self.ctx
        heat = np.asarray(heatmap)
        if heat.dtype not in (np.float32, np.float64):
            heat = heat.astype(np.float64)
        # steps 1-2: mesh into the depth camera's frame, then into the colour camera's frame (two float64 posings
        # like the reference's two mesh.transform calls would give one rounding more; the product is applied once)
        T_depth = np.linalg.inv(cur)
        T = np.linalg.inv(self.color_to_depth) @ T_depth
        ctx.pose_mesh(T)
        res = ctx.project(heat, self.K, None, self.threshold, frame="camera", accumulate=True,
                          want=("pixel", "t_hit", "face", "point64"))
        pix = res["pixel"].astype(np.int64)
        inten = heat.reshape(-1)[pix]
        # step 3: earlier clouds follow the object
        if self.previous_transformation is not None:
            rel = relative_transformation(cur, self.previous_transformation)
            for pcd in self.intersection_pcds:
                pcd.transform(rel)
        # step 4: selection + colours + colour->depth transform in one GPU pass
        pk = ctx.pack_hits(inten, res["face"], res["pixel"], res["point64"], T=self.color_to_depth,
                           want=("points", "colors", "face", "pixel"))
        pcd = _dp.PointCloud(pk["points"], pk["colors"])
        pcd.face_ids, pcd.pixels = pk["face"], pk["pixel"].astype(np.int64)
        self.intersection_pcds.append(pcd)
        self.previous_transformation = cur
        # the mesh the viewer gets is the caller's target_mesh_copy: the model posed by inv(current) ONLY, i.e. in the
        # DEPTH camera's frame -- the frame the clouds were just moved into (run.py:109-110 / :179-181 and :205);
        # the colour-camera copy that was traced stays inside the context
        self.mesh_in_camera = _dp.TriangleMesh(ctx.transform_points(np.asarray(self.V, dtype=np.float64), T_depth), self.F)
        return pcd

    def payload(self):
        """What update_dash_data ships (:205), plus the accumulated per-face arrays."""
        from .web_vis import build_payload
        p = build_payload(self.intersection_pcds, self.mesh_in_camera, with_face_intensity=False)
        p["face_hits"], p["face_intensity"], p["vertex_intensity"] = self.ctx.accum_get()
        return p
