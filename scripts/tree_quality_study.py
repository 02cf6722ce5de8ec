"""CPU study (numpy, no GPU): how many wide-node visits per ray would a better binary tree save?

Builds two binary hierarchies over the same synthetic mesh -- (A) the LBVH the library builds (30-bit Morton codes of
the box centres on per-axis extents, radix-tree topology) and (B) a top-down binned surface-area-heuristic tree -- puts
both through the SAME greedy 8-wide collapse (largest area first, leaf slots of <= 3 triangles) and the SAME ordered
closest-hit traversal, and counts the wide nodes fetched per ray on a sub-sampled dense WFOV frame.  The ratio B / A is
the head-room a better builder (treelet restructuring, PLOC, binned SAH on the GPU) could buy the traversal kernel,
whose time is proportional to node visits.  Usage: tree_quality_study.py [nu nv] (default 200 150 = 60k triangles)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import synth  # noqa: E402  (host-side generators only: no device needed)

LEAF_MAX = 3
sys.setrecursionlimit(100000)


# ------------------------------------------------------------------------------------------ binary builders
def morton30(c):
    q = np.clip((c * 1024.0).astype(np.int64), 0, 1023)

    def part(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x
    return (part(q[:, 0]) << 2) | (part(q[:, 1]) << 1) | part(q[:, 2])


class Tree:
    """binary tree: children arrays (negative = ~triangle index), boxes per internal node, triangle counts"""

    def __init__(self):
        self.left, self.right, self.lo, self.hi, self.count = [], [], [], [], []

    def add(self):
        for a in (self.left, self.right, self.lo, self.hi, self.count):
            a.append(None)
        return len(self.left) - 1


def build_lbvh(tlo, thi):
    cen = 0.5 * (tlo + thi)
    mn, mx = tlo.min(0), thi.max(0)
    codes = morton30((cen - mn) / np.maximum(mx - mn, 1e-30))
    order = np.argsort(codes, kind="stable")
    codes = codes[order]
    t = Tree()

    def rec(a, b):                                          # [a, b) of the sorted order, b - a >= 2
        me = t.add()
        if codes[a] == codes[b - 1]:
            m = (a + b) // 2                                # duplicate codes: split by index
        else:
            bit = int(codes[a] ^ codes[b - 1]).bit_length() - 1
            m = a + int(np.searchsorted(codes[a:b] >> bit, (codes[a] >> bit) + 1))
        kids = []
        for (x, y) in ((a, m), (m, b)):
            kids.append(~int(order[x]) if y - x == 1 else rec(x, y))
        t.left[me], t.right[me] = kids
        idx = order[a:b]
        t.lo[me], t.hi[me], t.count[me] = tlo[idx].min(0), thi[idx].max(0), b - a
        return me
    rec(0, len(order))
    return t


def build_sah(tlo, thi, bins=16):
    cen = 0.5 * (tlo + thi)
    t = Tree()

    def area(lo, hi):
        d = hi - lo
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]

    def rec(idx):
        me = t.add()
        lo, hi = tlo[idx].min(0), thi[idx].max(0)
        t.lo[me], t.hi[me], t.count[me] = lo, hi, len(idx)
        best = (np.inf, None)
        clo, chi = cen[idx].min(0), cen[idx].max(0)
        for ax in range(3):
            ext = chi[ax] - clo[ax]
            if ext <= 0:
                continue
            b = np.minimum(((cen[idx, ax] - clo[ax]) / ext * bins).astype(np.int64), bins - 1)
            for s in range(1, bins):
                L = idx[b < s]
                R = idx[b >= s]
                if len(L) == 0 or len(R) == 0:
                    continue
                c = area(tlo[L].min(0), thi[L].max(0)) * len(L) + area(tlo[R].min(0), thi[R].max(0)) * len(R)
                if c < best[0]:
                    best = (c, (L, R))
        if best[1] is None:                                 # all centroids equal: split by index
            h = len(idx) // 2
            best = (0.0, (idx[:h], idx[h:]))
        kids = [~int(part[0]) if len(part) == 1 else rec(part) for part in best[1]]
        t.left[me], t.right[me] = kids
        return me
    rec(np.arange(len(tlo)))
    return t


# ------------------------------------------------------------------------------------------ 8-wide collapse
def collapse(t, tlo, thi):
    """wide nodes: list of (child boxes lo [k,3], hi [k,3], kinds [k] (wide node index or -1), leaf triangle lists)"""
    def box(c):
        return (tlo[~c], thi[~c]) if c < 0 else (t.lo[c], t.hi[c])

    def cnt(c):
        return 1 if c < 0 else t.count[c]

    def area(c):
        lo, hi = box(c)
        d = hi - lo
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]

    def tris(c, out):
        if c < 0:
            out.append(~c)
        else:
            tris(t.left[c], out)
            tris(t.right[c], out)
        return out
    wide = []

    def make(root):
        me = len(wide)
        wide.append(None)
        cand = [t.left[root], t.right[root]]
        while len(cand) < 8:
            exp = [(area(c), i) for i, c in enumerate(cand) if cnt(c) > LEAF_MAX]
            if not exp:
                break
            _, i = max(exp)
            c = cand[i]
            cand[i] = t.left[c]
            cand.append(t.right[c])
        los, his, kinds, leaves = [], [], [], []
        for c in cand:
            lo, hi = box(c)
            los.append(lo)
            his.append(hi)
            if cnt(c) > LEAF_MAX:
                kinds.append(make(c))
                leaves.append(None)
            else:
                kinds.append(-1)
                leaves.append(tris(c, []))
        wide[me] = (np.array(los, np.float64), np.array(his, np.float64), kinds, leaves)
        return me
    make(0)
    return wide


# ------------------------------------------------------------------------------------------ traversal
def trace(wide, V, F, rays6):
    v0, e1, e2 = V[F[:, 0]], V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]
    nodes = tris = 0
    faces = np.full(len(rays6), -1, np.int64)
    for r, ray in enumerate(rays6):
        o, d = ray[:3].astype(np.float64), ray[3:].astype(np.float64)
        inv = 1.0 / np.where(d == 0.0, 1e-300, d)
        best, bf = np.inf, -1
        stack = [(0.0, 0)]
        while stack:
            tent, w = stack.pop()
            if tent > best:
                continue
            nodes += 1
            lo, hi, kinds, leaves = wide[w]
            t0, t1 = (lo - o) * inv, (hi - o) * inv
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            hit = (np.maximum(tn, 0.0) <= np.minimum(tf, best))
            inner = []
            for k in np.nonzero(hit)[0]:
                if kinds[k] >= 0:
                    inner.append((max(tn[k], 0.0), kinds[k]))
                else:
                    for f in leaves[k]:
                        tris += 1
                        p = np.cross(d, e2[f])
                        det = e1[f] @ p
                        if det == 0.0:
                            continue
                        s = o - v0[f]
                        u = (s @ p) / det
                        q = np.cross(s, e1[f])
                        v = (d @ q) / det
                        tt = (e2[f] @ q) / det
                        if u >= 0 and v >= 0 and u + v <= 1 and tt >= 0 and (tt < best or (tt == best and f < bf)):
                            best, bf = tt, f
            for item in sorted(inner, reverse=True):        # nearest popped first
                stack.append(item)
        faces[r] = bf
    return nodes / len(rays6), tris / len(rays6), faces


def sah_cost(wide):
    lo0 = np.minimum.reduce([w[0].min(0) for w in wide[:1]])
    hi0 = np.maximum.reduce([w[1].max(0) for w in wide[:1]])
    d = hi0 - lo0
    a_root = d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
    c = 1.0                                                 # the root is always visited
    for lo, hi, kinds, _ in wide:
        dd = hi - lo
        a = dd[:, 0] * dd[:, 1] + dd[:, 1] * dd[:, 2] + dd[:, 2] * dd[:, 0]
        c += sum(a[k] for k in range(len(kinds)) if kinds[k] >= 0) / a_root
    return c


def main():
    nu, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 150)
    V, F = synth.param_mesh(nu, nv, seed=0, scale=6.0)
    V = V.astype(np.float64)
    tlo, thi = V[F].min(1), V[F].max(1)
    K, H, W = synth.camera_wfov()
    pose = synth.fill_frame_pose()
    step = 8
    ys, xs = np.meshgrid(np.arange(step // 2, H, step), np.arange(step // 2, W, step), indexing="ij")
    xs, ys = xs.ravel().astype(np.float64), ys.ravel().astype(np.float64)
    dcam = np.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], np.ones_like(xs)], 1)
    dcam /= np.linalg.norm(dcam, axis=1, keepdims=True)
    Ri, ti = pose[:3, :3].T, -pose[:3, :3].T @ pose[:3, 3]
    rays6 = np.concatenate([np.tile(ti, (len(xs), 1)), dcam @ Ri.T], 1)
    out = {"triangles": len(F), "rays": len(rays6)}
    res = {}
    for name, builder in (("lbvh", build_lbvh), ("sah", build_sah)):
        t0 = time.perf_counter()
        tree = builder(tlo, thi)
        wide = collapse(tree, tlo, thi)
        t1 = time.perf_counter()
        n, tr, faces = trace(wide, V, F, rays6)
        res[name] = faces
        out[name] = {"wide_nodes": len(wide), "expected_inner_visits_random_rays": sah_cost(wide),
                     "nodes_per_ray": n, "tris_per_ray": tr, "hit_frac": float((faces >= 0).mean()),
                     "build_s": t1 - t0, "trace_s": time.perf_counter() - t1}
        print(name, json.dumps(out[name]), flush=True)
    out["same_faces"] = bool(np.array_equal(res["lbvh"], res["sah"]))
    out["nodes_per_ray_ratio_sah_over_lbvh"] = out["sah"]["nodes_per_ray"] / out["lbvh"]["nodes_per_ray"]
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r1d_tree_quality_study_%dx%d.json" % (nu, nv)), "w"), indent=1)


if __name__ == "__main__":
    main()
