"""CPU study (numpy): leaf slots of at most 1/2/3/4/6/8 triangles under the same LBVH, collapse and traversal; cost model
12 x nodes + 3.5 x triangles per ray (warp instructions per active lane: a 308-instruction visit at 25.5 lanes, a 32-pair
triangle batch of ~110 instructions)."""
import sys, json, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, '6dof-pose-estimation-and-defect-projection_b200'))
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
src = open(os.path.join(ROOT, 'scripts', 'tree_quality_study.py')).read().replace('if __name__ == "__main__":\n    main()', '')
base = {'__file__': os.path.join(ROOT, 'scripts', 'tree_quality_study.py'), '__name__': 'tq_w'}
ns0 = dict(base); exec(compile(src, 'tq', 'exec'), ns0)
tree = ns0['build_lbvh'](tlo, thi)
for leaf in (1, 2, 3, 4, 6, 8):
    ns = dict(base); exec(compile(src.replace("LEAF_MAX = 3", f"LEAF_MAX = {leaf}"), 'tq', 'exec'), ns)
    wide = ns['collapse'](tree, tlo, thi)
    n, tr, faces = ns['trace'](wide, V, F, rays6)
    print(json.dumps({"leaf_max": leaf, "wide_nodes": len(wide), "nodes_per_ray": n, "tris_per_ray": tr, "model_cost_12n_plus_3p5t": 12 * n + 3.5 * tr}), flush=True)
