#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python.

Run in the authoring container only (needs /root/reference); the fixtures it
writes are committed and are what travels to the GPU box.

    python tests/golden/make_golden.py

How: /root/reference/src/defect_projection.py is executed unmodified, with
`open3d` and `matplotlib` replaced by the minimal shims in ref_shims.py
(both packages are absent offline).  Everything the reference computes itself
in numpy -- heatmap_to_points (:165-179), compute_rays (:196-223), the
float32 ray tensor and hit-point formula of intersect_rays_with_mesh
(:245-264), create_intersection_pcd (:268-294), project_debug_rays (:296-317),
generate_centered_heatmap (:137-155, real cv2), load_extrinsics (:65-92) and the
ray_tracing orchestration (:527-563) -- therefore runs as written.  The single
substituted piece is RaycastingScene.cast_rays (Embree): the shim answers it
with oracle semantic (A) (float32 watertight closest hit, oracle/oracle.c).
That boundary stays "parity unpinned" (see oracle/oracle.c header).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
from defectproj import synth  # noqa: E402

REF = "/root/reference/src/defect_projection.py"


def load_reference():
    ref_shims.install()
    ns = {"__name__": "reference_defect_projection"}
    with open(REF) as f:
        src = f.read()
    exec(compile(src, REF, "exec"), ns)
    return ns


def main():
    ref = load_reference()
    o3d = sys.modules["open3d"]
    out = {}

    # ---- G1: heatmap_to_points ordering example + reference Gaussian heatmap (real cv2)
    h34 = np.array([[0.1, 0.2, 0.3, 0.9],
                    [0.8, 0.5, 0.0, 0.2],
                    [0.0, 0.51, 0.5, 0.4]], dtype=np.float64)
    pts = ref["heatmap_to_points"](h34, 0.5)
    out["g1_heat"] = h34
    out["g1_points"] = np.array([[p[0], p[1], p[2]] for p in pts], dtype=np.float64)

    hg = ref["generate_centered_heatmap"]((96, 128), 1.0, 12)
    out["g2_heat"] = hg
    for thr, key in ((0.5, "g2_points_050"), (0.75, "g2_points_075")):
        pts = ref["heatmap_to_points"](hg, thr)
        out[key] = np.array([[p[0], p[1], p[2]] for p in pts], dtype=np.float64)

    # counts of the production-size heatmap: stored as numbers only
    hbig = ref["generate_centered_heatmap"]((720, 1280), 1.0, 50)
    out["g3_counts"] = np.array([len(ref["heatmap_to_points"](hbig, 0.5)),
                                 len(ref["heatmap_to_points"](hbig, 0.75))], dtype=np.int64)

    # ---- G4: compute_rays (float64) on the g2 points with a scaled-down 720p camera
    K = synth.K_matrix(61.0, 61.5, 64.0, 48.0)
    intr = o3d.camera.PinholeCameraIntrinsic(128, 96, K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    pts = ref["heatmap_to_points"](hg, 0.5)
    rays, inten = ref["compute_rays"](pts, intr)
    out["g4_K"] = K
    out["g4_rays"] = rays
    out["g4_intensities"] = inten

    # ---- G5: the whole ray_tracing() call on a small synthetic scene
    V, F = synth.param_mesh(24, 16, seed=3)
    T_icp_inv = synth.fixed_pose(z=420.0)               # what run.py:109-110 applies first
    c2d = np.eye(4)
    c2d[:3, :3] = synth.rot_y(1.5) @ synth.rot_x(-0.7)
    c2d[:3, 3] = [-32.0, -2.0, 4.0]                     # Azure-Kinect-like colour->depth offset (mm)
    mesh = o3d.geometry.TriangleMesh(V.astype(np.float64), F)
    mesh.transform(T_icp_inv)                           # mesh in the depth-camera frame, mm
    out["g5_V_model"] = V
    out["g5_F"] = F
    out["g5_V_depthcam"] = np.asarray(mesh.vertices).copy()
    out["g5_color_to_depth"] = c2d
    out["g5_K"] = K
    with tempfile.TemporaryDirectory() as d:
        synth.write_scene_dir(d, K, (96, 128), color_to_depth=c2d)
        for thr, tag in ((0.5, "050"), (0.75, "075")):
            pcd, mesh_out = ref["ray_tracing"](d, mesh, hg, intr, heatmap_threshold=thr)
            out[f"g5_points_{tag}"] = np.asarray(pcd.points).copy()
            out[f"g5_colors_{tag}"] = np.asarray(pcd.colors).copy()
            out[f"g5_V_colorcam_{tag}"] = np.asarray(mesh_out.vertices).copy()
        # miss-all case: aim the heatmap blob where the mesh is not -> LineSet branch (:561-563)
        far = o3d.geometry.TriangleMesh(V.astype(np.float64) + np.array([5000.0, 0, 0]), F)
        ls, _ = ref["ray_tracing"](d, far, hg, intr, heatmap_threshold=0.9)
        out["g6_lineset_points"] = np.asarray(ls.points).copy()
        out["g6_lineset_lines"] = np.asarray(ls.lines).copy()
        out["g6_lineset_colors"] = np.asarray(ls.colors).copy()

    # ---- G7: jet colour mapping of create_intersection_pcd on a ramp and on a constant
    ramp = np.linspace(0.2, 0.9, 33)
    pc = ref["create_intersection_pcd"](np.zeros((33, 3)), ramp)
    out["g7_ramp"] = ramp
    out["g7_colors"] = np.asarray(pc.colors).copy()

    path = os.path.join(HERE, "reference_path.npz")
    np.savez_compressed(path, **out)

    # ---- D: the depth-image projection path (:359-492, :613-630), file depth_path.npz
    dep = {}
    rng = np.random.default_rng(7)
    Hh, Ww = 60, 80
    heat = (synth.gaussian_heatmap((Hh, Ww), sigma=11.0) + 0.3 * synth.blob_heatmap((Hh, Ww), seed=5, dtype=np.float64)
            * (rng.random((Hh, Ww)) < 0.5)) * 3.7                                # max != 1: the /max matters
    depth = rng.integers(300, 900, (Hh, Ww)).astype(np.uint16)
    depth[rng.random((Hh, Ww)) < 0.2] = 0                                       # invalid depth pixels
    Kd = synth.K_matrix(75.3, 76.1, 39.5, 30.25)
    intr_d = o3d.camera.PinholeCameraIntrinsic(Ww, Hh, Kd[0, 0], Kd[1, 1], Kd[0, 2], Kd[1, 2])
    dep["d_heat"], dep["d_depth"], dep["d_K"] = heat, depth, Kd
    for thr, tag in ((0.1, "010"), (0.5, "050")):
        dep[f"d_point3d_{tag}"] = ref["heatmap_to_point3d"](heat, depth, intr_d, threshold=thr)
    # a float32 map: the quotient heatmap[y, x] / max_value is a float32 division (:384) whose float32-rounded value lands
    # in column 3.  The threshold comparison np.float32 > python float is float64 under the reference's numpy 1.26.4 and
    # float32 under numpy >= 2 (NEP 50); the fixture must not depend on that, so no quotient may sit between the two
    heat32 = heat.astype(np.float32)
    q32 = heat32 / np.max(heat32)
    for thr, tag in ((0.1, "010"), (0.5, "050")):
        assert not ((q32.astype(np.float64) > thr) ^ (q32 > np.float32(thr))).any()
        dep[f"d_point3d_f32_{tag}"] = ref["heatmap_to_point3d"](heat32, depth, intr_d, threshold=thr)
    small_depth = depth[:50, :70]                                               # depth image smaller than the heatmap
    dep["d_point3d_small"] = ref["heatmap_to_point3d"](heat, small_depth, intr_d, threshold=0.3)
    picks = [(int(x), int(y)) for x, y in zip(rng.integers(0, Ww, 40), rng.integers(0, Hh, 40))]
    dep["d_picks"] = np.array(picks, np.int32)
    dep["d_calc"] = ref["calc_coordinates"](depth, picks, intr_d)
    # target cloud with normals: the vertices of a small synthetic mesh, normals = normalised radial direction
    Vt, Ft = synth.param_mesh(30, 20, seed=9)
    tp = Vt.astype(np.float64) * 2.0 + np.array([0.0, 0.0, 600.0])
    tn = tp - tp.mean(0)
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    target = o3d.geometry.PointCloud()
    target.points = o3d.utility.Vector3dVector(tp)
    target.normals = o3d.utility.Vector3dVector(tn)
    dep["d_target_points"], dep["d_target_normals"] = tp, tn
    offs, ali, p3 = ref["depth_projection_heatmap"](depth, intr_d, target, heat)
    dep["d_proj_offset"], dep["d_proj_aligned"], dep["d_proj_point3d"] = offs, ali, p3
    o2, a2 = ref["align_to_surface"](dep["d_point3d_050"], target, offset=0.1)
    dep["d_align_offset_01"], dep["d_align_aligned_01"] = o2, a2
    # ---- E: DataReader.get_heatmap (datareader.py:639-675) executed verbatim with the real cv2
    import ast
    import textwrap
    import types as _types
    import cv2
    src_dr = open("/root/reference/datareader.py").read()
    tree = ast.parse(src_dr)
    fn_src = None
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "DataReader":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "get_heatmap":
                    fn_src = ast.get_source_segment(src_dr, item)
    ns2 = {"np": np, "cv2": cv2}
    exec(textwrap.dedent(fn_src), ns2)
    for tag, (hs, cH, cW, ds) in {"a": (64, 180, 320, 1), "b": (100, 90, 76, 1), "c": (224, 360, 640, 2),
                                  "d": (96, 200, 150, 1)}.items():
        data = (rng.random((hs, hs)) * 5.0 - 1.0) * synth.gaussian_heatmap((hs, hs), sigma=hs / 5.0)
        if tag == "d":
            data = data.astype(np.float32)      # a float32 .npy stays float32 through :658-665 (CV_32F resize)
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "heatmap"))
            np.save(os.path.join(d, "heatmap", "0002.npy"), data)
            fake = _types.SimpleNamespace(base_dir=d, color_H=cH, color_W=cW, downscale=ds)
            color = (rng.random((cH, cW, 3)) * 255).astype(np.uint8)
            full, _, vis, _ = ns2["get_heatmap"](fake, color)
        dep[f"h_data_{tag}"], dep[f"h_full_{tag}"] = data, full
        dep[f"h_cfg_{tag}"] = np.array([cH, cW, ds], np.int64)
    np.savez_compressed(os.path.join(HERE, "depth_path.npz"), **dep)
    print("wrote depth_path.npz", {k: v.shape for k, v in dep.items()})
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("g3 counts (720p sigma=50, thr .5/.75):", out["g3_counts"])


if __name__ == "__main__":
    main()
