"""Viewer payload: the step after the back-projection.

``update_dash_data`` mirrors /root/reference/src/web_vis.py:203-217: it builds the dictionary the Dash thread
takes from its Queue -- ``{'pcds': [{'points', 'colors'}, ...], 'vertices', 'faces'}`` -- so ``update_figure``
(:147-168) works unchanged.  Here the arrays come straight from the GPU results (hit selection, jet colours and
the colour->depth transform are one kernel pass, ``dp_pack_hits``), and the payload additionally carries the
per-face / per-vertex defect intensities the accumulators hold, which is what
``go.Mesh3d(intensity=..., intensitymode='cell')`` needs to paint the defect ON the mesh instead of as a cloud.
No Dash / plotly import happens here: the queue is any object with ``put``.
"""
from __future__ import annotations

import numpy as np

from . import defect_projection as _dp

__all__ = ["build_payload", "update_dash_data", "mesh3d_kwargs"]

data_queue = None


def build_payload(intersection_pcds, target_mesh, with_face_intensity=True):
    pcd_data = []
    for pcd in intersection_pcds:
        pcd_data.append({"points": np.asarray(pcd.points), "colors": np.asarray(pcd.colors)})
    payload = {"pcds": pcd_data,
               "vertices": np.asarray(target_mesh.vertices),
               "faces": np.asarray(target_mesh.triangles)}
    if with_face_intensity:
        hist, fmax, vmax = _dp.face_intensities()
        if hist is not None and len(hist) == len(payload["faces"]):
            payload["face_hits"] = hist
            payload["face_intensity"] = fmax
            payload["vertex_intensity"] = vmax
    return payload


def update_dash_data(intersection_pcds, target_mesh, queue=None):
    """Same call as the reference's; ``queue`` defaults to the module-level ``data_queue`` (set by the app)."""
    payload = build_payload(intersection_pcds, target_mesh)
    q = queue if queue is not None else data_queue
    if q is not None:
        q.put(payload)
    return payload


def mesh3d_kwargs(payload, mode="cell"):
    """Keyword arguments for plotly's go.Mesh3d that paint the accumulated defect intensity on the mesh
    (``intensitymode='cell'``: one value per face; ``'vertex'``: one per vertex)."""
    v, f = payload["vertices"], payload["faces"]
    kw = dict(x=v[:, 0], y=v[:, 1], z=v[:, 2], i=f[:, 0], j=f[:, 1], k=f[:, 2], opacity=1)
    key = "face_intensity" if mode == "cell" else "vertex_intensity"
    if key in payload:
        kw.update(intensity=payload[key], intensitymode=mode, colorscale="Jet", cmin=0.0, cmax=1.0)
    else:
        kw.update(color="grey")
    return kw
