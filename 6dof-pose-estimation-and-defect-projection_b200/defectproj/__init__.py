"""defectproj -- B200-native 2D-defect -> 3D-mesh back-projection (drop-in for the hot path of
ziadabohalawa/6DoF-Pose-Estimation-and-Defect-Projection, src/defect_projection.py).

    from defectproj import defect_projection as dp      # the reference's call surface
    from defectproj import Context, Projector            # the batched / multi-GPU surface
    from defectproj import datareader, web_vis, tracking # the steps either side of the path (get_heatmap,
                                                         # update_dash_data, the run.py detection loop)

All computation runs in libdefectproj.so (hand-written sm_100a CUDA behind a C ABI,
include/defectproj.h).  There is no CPU fallback.
"""
from . import synth  # noqa: F401
from .core import Context, DefectProjError  # noqa: F401
from .projector import FrameStream, Projector  # noqa: F401

__all__ = ["Context", "Projector", "FrameStream", "DefectProjError", "synth"]
