"""ctypes binding of libdefectproj.so (include/defectproj.h).

The library is the product: there is no Python or CPU fallback.  Importing this
module without the built .so raises; creating a context without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("DEFECTPROJ_LIB") or os.path.join(_HERE, "libdefectproj.so")

DP_OK, DP_E_ARG, DP_E_CUDA, DP_E_NOMEM, DP_E_STATE = 0, -1, -2, -3, -4
DP_HOST, DP_DEVICE = 0, 1
DP_F32, DP_F64 = 0, 1
DP_FRAME_OBJECT, DP_FRAME_CAMERA = 0, 1
DP_PEER_HANDLE_BYTES = 64
DP_PEER_STAGE, DP_PEER_RECORDS, DP_PEER_REC_COUNT, DP_PEER_T_HIT, DP_PEER_FACE, DP_PEER_POINT = range(6)
ABI_VERSION = 1

i64, i32, f64, vp = C.c_int64, C.c_int, C.c_double, C.c_void_p


class RaysOut(C.Structure):
    _fields_ = [("pixel", vp), ("intensity", vp), ("t_hit", vp), ("face", vp), ("point", vp), ("point64", vp),
                ("cap", i64), ("counts", vp)]


class Stats(C.Structure):
    _fields_ = [("rays", i64), ("hits", i64), ("nodes_fetched", i64), ("tris_tested", i64),
                ("n_wide_nodes", i64), ("n_tris", i64), ("wide_depth", C.c_int32), ("reserved", C.c_int32),
                ("last_build_ms", C.c_float), ("last_refit_ms", C.c_float)]


# name -> (restype, argtypes); every symbol include/defectproj.h declares
SYMBOLS = {
    "dp_abi_version": (i32, []),
    "dp_create": (i32, [i32, C.POINTER(vp)]),
    "dp_destroy": (None, [vp]),
    "dp_last_error": (C.c_char_p, [vp]),
    "dp_synchronize": (i32, [vp, vp]),
    "dp_set_mesh": (i32, [vp, vp, i32, i64, vp, i64, i32, vp]),
    "dp_build_bvh": (i32, [vp, vp]),
    "dp_update_vertices": (i32, [vp, vp, i32, i64, i32, vp]),
    "dp_pose_mesh": (i32, [vp, vp, vp]),
    "dp_get_posed_vertices": (i32, [vp, vp, i32, i32, vp]),
    "dp_compact": (i32, [vp, vp, i32, i64, i32, i32, f64, vp, vp, i64, C.POINTER(i64), vp, i32, vp]),
    "dp_frame_xform": (None, [vp, vp, vp]),
    "dp_compute_rays": (i32, [vp, vp, vp, i64, vp, vp, i32, vp]),
    "dp_cast_rays": (i32, [vp, i32, vp, i64, vp, vp, i32, vp]),
    "dp_project": (i32, [vp, i32, vp, i32, i64, i32, i32, f64, vp, i64, vp, i32, C.POINTER(RaysOut),
                         C.POINTER(i64), C.POINTER(i64), i32, vp]),
    "dp_depth_backproject": (i32, [vp, vp, i32, i32, i32, vp, i32, i32, vp, f64, vp, i64, C.POINTER(i64), i32, vp]),
    "dp_calc_coordinates": (i32, [vp, vp, vp, i64, vp, i32, i32, vp, vp, vp, i32, vp]),
    "dp_align_to_surface": (i32, [vp, vp, i32, i64, vp, vp, i64, f64, vp, vp, vp, i32, vp]),
    "dp_estimate_normals": (i32, [vp, vp, i64, f64, i32, vp, i32, vp, i32, vp]),
    "dp_prepare_heatmap": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp]),
    "dp_transform_points": (i32, [vp, vp, i64, vp, i32, vp]),
    "dp_pack_hits": (i32, [vp, vp, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, C.POINTER(i64), i32, vp]),
    "dp_jet_lut": (None, [vp]),
    "dp_set_ray_shard": (i32, [vp, i32, i32]),
    "dp_shard_slots": (i32, [i32, i32, i64, i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]),
    "dp_pack_records": (i32, [vp, vp, vp, vp, vp, i64, i64, vp, i64, C.POINTER(i64), vp, i32, vp]),
    "dp_peer_export": (i32, [vp, i64, i64, vp, C.POINTER(i64)]),
    "dp_peer_open": (i32, [vp, i32, i32, vp]),
    "dp_peer_open_local": (i32, [vp, i32, i32, C.POINTER(vp)]),
    "dp_peer_close": (i32, [vp]),
    "dp_peer_window": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(i64)]),
    "dp_peer_snapshot": (i32, [vp, i32, i32, vp]),
    "dp_peer_combine": (i32, [vp, i32, vp, i32, vp, i64, i32, vp, vp]),
    "dp_peer_results": (i32, [vp, i32, i32]),
    "dp_peer_status": (i32, [vp, C.POINTER(i32)]),
    "dp_accum_layout": (i32, [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "dp_icp_point_to_plane": (i32, [vp, vp, i64, vp, vp, i64, f64, vp, i32, f64, f64, vp, C.POINTER(f64), C.POINTER(f64),
                                    C.POINTER(i32), vp, i32, vp]),
    "dp_accum_reset": (i32, [vp, vp]),
    "dp_accum_flush": (i32, [vp, vp]),
    "dp_accum_get": (i32, [vp, vp, vp, vp, i32, vp]),
    "dp_accum_device_ptrs": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "dp_set_stats": (i32, [vp, i32]),
    "dp_get_stats": (i32, [vp, C.POINTER(Stats)]),
    "dp_last_timings": (i32, [vp, vp]),
    "dp_set_timing": (i32, [vp, i32]),
    "dp_debug_dump_bvh": (i32, [vp, i32, vp, C.POINTER(i64), vp, C.POINTER(i64)]),
    "dp_debug_ray_nodes": (i32, [vp, vp, i64]),
    "dp_debug_radix_sort": (i32, [vp, vp, vp, i64]),
    "dp_debug_morton": (i32, [vp, vp]),
}

_lib = None


def load():
    """Load the C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  defectproj has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.dp_abi_version() != ABI_VERSION:
        raise ImportError(f"libdefectproj.so ABI {L.dp_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L
