import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth
from oracle import oracle as orc
V, F = synth.param_mesh(40, 25, seed=2)
ctx = Context(0); ctx.set_mesh(V, F).build_bvh()
rng = np.random.default_rng(3)
o = np.array([0.0, 0.0, 300.0], np.float32)
d = V[rng.integers(0, len(V), 3000)] - o
rays6 = np.ascontiguousarray(np.hstack([np.tile(o, (len(d), 1)), d]), np.float32)
t, f = ctx.cast_rays(rays6)
t0, f0 = orc.cast_brute_f32(V, F, rays6)
bad = np.nonzero((f != f0) | (t.view(np.uint32) != t0.view(np.uint32)))[0]
print("lib", os.environ.get("DEFECTPROJ_LIB", "default"), "refill", os.environ.get("DP_REFILL"), "bad:", bad, f[bad], f0[bad], t[bad], t0[bad])
for i in bad[:3]:
    # every face the oracle says this ray hits
    r = rays6[i]
    hits = []
    for ff in range(len(F)):
        h, tt = orc.tri_test_f32(r, V[F[ff, 0]], V[F[ff, 1]], V[F[ff, 2]])
        if h: hits.append((ff, np.float32(tt)))
    print("ray", i, r, "oracle candidates:", hits)
    t1, f1 = ctx.cast_rays(rays6[i:i+1]); print("  alone:", t1, f1)
