"""Heatmap preparation: the step before the back-projection.

Mirrors ``DataReader.get_heatmap`` of /root/reference/datareader.py:639-675.  The array work on the path
(min-max normalisation :658-659, the cv2 INTER_LINEAR resize :664-665, the zero padding :669-674) runs in
libdefectproj.so (``dp_prepare_heatmap``, csrc/prep.cu) and is bit-exact with the reference's cv2 output for
float32 and float64 maps.  The colour-image crops the reference returns alongside (:649-656, :666-667) only
feed the 2-D overlay PNG, not the projection; they are produced with cv2 on the host when cv2 is importable and
are ``None`` otherwise.
"""
from __future__ import annotations

import numpy as np

from .defect_projection import get_context

__all__ = ["prepare_heatmap", "HeatmapReader"]


def prepare_heatmap(heatmap_data, color_H, color_W, downscale=1, out_dtype=np.float64, with_vis=False):
    """``heatmap_full`` of get_heatmap: [int(color_H/downscale), int(color_W/downscale)] with the normalised map,
    resized to the shorter side, in the centre.  ``with_vis`` also returns the o x o window (``heatmap_vis``)."""
    H, W = int(color_H / downscale), int(color_W / downscale)
    full = get_context().prepare_heatmap(heatmap_data, H, W, out_dtype)
    if not with_vis:
        return full
    o = min(H, W)
    y0, x0 = (H - o) // 2, (W - o) // 2
    vis = full[y0:y0 + o, x0:x0 + o]
    # cv2.resize keeps the map's dtype (:664); heatmap_full is float64 (:669)
    src_dt = getattr(heatmap_data, "dtype", np.float64)
    return full, vis.astype(np.float32) if str(src_dt).endswith("float32") else vis


class HeatmapReader:
    """The slice of DataReader that get_heatmap touches: ``base_dir``, ``color_H``, ``color_W``, ``downscale``."""

    def __init__(self, base_dir, color_H, color_W, downscale=1):
        self.base_dir, self.color_H, self.color_W, self.downscale = base_dir, int(color_H), int(color_W), downscale

    def get_heatmap(self, color_image=None, device=False):
        """(heatmap_full, color_original, heatmap_vis, color_original) like datareader.py:675.  device=True: the raw map
        (224x224: 400 KB) is uploaded and heatmap_full / heatmap_vis are CUDA tensors -- ray_tracing() takes heatmap_full as
        it is, so the padded 720p map (7.4 MB as float64) never crosses PCIe."""
        heatmap_data = np.load(f"{self.base_dir}/heatmap/0002.npy")
        if device:
            import torch
            ctx = get_context()
            data = heatmap_data if heatmap_data.dtype in (np.float32, np.float64) else heatmap_data.astype(np.float64)
            data_t = torch.from_numpy(np.ascontiguousarray(data)).to(f"cuda:{ctx.device}")
            H, W = int(self.color_H / self.downscale), int(self.color_W / self.downscale)
            full = ctx.prepare_heatmap(data_t, H, W, np.float64)
            o = min(H, W)
            y0, x0 = (H - o) // 2, (W - o) // 2
            vis = full[y0:y0 + o, x0:x0 + o]
            if data.dtype == np.float32:
                vis = vis.float()
        else:
            full, vis = prepare_heatmap(heatmap_data, self.color_H, self.color_W, self.downscale, with_vis=True)
        color_original = None
        if color_image is not None:
            try:
                import cv2
            except ImportError:
                cv2 = None
            if cv2 is not None:
                hs = heatmap_data.shape[0]
                scale = hs / min(color_image.shape[:2])
                nh, nw = int(color_image.shape[0] * scale), int(color_image.shape[1] * scale)
                resized = cv2.resize(color_image, (nw, nh), interpolation=cv2.INTER_AREA)
                sy, sx = (nh - hs) // 2, (nw - hs) // 2
                o = vis.shape[0]
                color_original = cv2.resize(resized[sy:sy + hs, sx:sx + hs], (o, o), interpolation=cv2.INTER_NEAREST)
        return full, color_original, vis, color_original
