"""CPU study (numpy): wide nodes fetched per ray for 4/6/8/12/16-wide collapses of the same LBVH (60k-triangle torus, dense
WFOV frame), and the instruction model visits x (140 + 21 x width) read off the SASS of the node visit."""
import sys, json, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '6dof-pose-estimation-and-defect-projection_b200'))
import tree_quality_study as tq
from defectproj import synth
nu, nv = 200, 150
V, F = synth.param_mesh(nu, nv, seed=0, scale=6.0); V = V.astype(np.float64)
tlo, thi = V[F].min(1), V[F].max(1)
K, H, W = synth.camera_wfov(); pose = synth.fill_frame_pose()
step = 8
ys, xs = np.meshgrid(np.arange(step // 2, H, step), np.arange(step // 2, W, step), indexing="ij")
xs, ys = xs.ravel().astype(np.float64), ys.ravel().astype(np.float64)
dcam = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs)], 1)
dcam /= np.linalg.norm(dcam, axis=1, keepdims=True)
Ri, ti = pose[:3, :3].T, -pose[:3, :3].T @ pose[:3, 3]
rays6 = np.concatenate([np.tile(ti, (len(xs), 1)), dcam @ Ri.T], 1)[::4]
tree = tq.build_lbvh(tlo, thi)
import types
src = open(os.path.join(ROOT, 'scripts', 'tree_quality_study.py')).read()
for width in (4, 6, 8, 12, 16):
    ns = {'__file__': os.path.join(ROOT, 'scripts', 'tree_quality_study.py'), '__name__': 'tq_w'}
    exec(compile(src.replace("while len(cand) < 8:", f"while len(cand) < {width}:").replace('if __name__ == "__main__":\n    main()', ''), 'tq_w', 'exec'), ns)
    wide = ns['collapse'](tree, tlo, thi)
    n, tr, faces = ns['trace'](wide, V, F, rays6)
    print(json.dumps({"width": width, "wide_nodes": len(wide), "nodes_per_ray": n, "tris_per_ray": tr, "model_cost_140_plus_21w": n * (140 + 21 * width)}), flush=True)
