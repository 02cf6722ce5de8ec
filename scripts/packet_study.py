"""CPU study (numpy): lock-step packets pay the longest ray of their 32.  Per-ray wide-node visit counts of a full-resolution
crop of the dense WFOV frame (LBVH, greedy 8-wide collapse, ordered closest-hit traversal) and the lane efficiency
mean(visits) / mean(max over the packet) for different shapes of the 32-ray packet."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
import tree_quality_study as tq  # noqa: E402
from defectproj import synth  # noqa: E402


def visits_per_ray(wide, rays6):
    out = np.zeros(len(rays6), np.int32)
    for r, ray in enumerate(rays6):
        o, d = ray[:3], ray[3:]
        inv = 1.0 / np.where(d == 0.0, 1e-300, d)
        best = np.inf
        stack = [(0.0, 0)]
        n = 0
        while stack:
            tent, w = stack.pop()
            if tent > best:
                continue
            n += 1
            lo, hi, kinds, leaves = wide[w]
            t0, t1 = (lo - o) * inv, (hi - o) * inv
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            hit = np.maximum(tn, 0.0) <= np.minimum(tf, best)
            inner = []
            for k in np.nonzero(hit)[0]:
                if kinds[k] >= 0:
                    inner.append((max(tn[k], 0.0), kinds[k]))
                else:
                    for f in leaves[k]:
                        tt = TRI(o, d, f)
                        if tt < best:
                            best = tt
            for item in sorted(inner, reverse=True):
                stack.append(item)
        out[r] = n
    return out


def main():
    global TRI
    V, F = synth.param_mesh(200, 150, seed=0, scale=6.0)
    V = V.astype(np.float64)
    tlo, thi = V[F].min(1), V[F].max(1)
    v0, e1, e2 = V[F[:, 0]], V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]

    def tri(o, d, f):
        p = np.cross(d, e2[f])
        det = e1[f] @ p
        if det == 0.0:
            return np.inf
        s = o - v0[f]
        u = (s @ p) / det
        q = np.cross(s, e1[f])
        v = (d @ q) / det
        tt = (e2[f] @ q) / det
        return tt if (u >= 0 and v >= 0 and u + v <= 1 and tt >= 0) else np.inf
    TRI = tri
    wide = tq.collapse(tq.build_lbvh(tlo, thi), tlo, thi)
    K, H, W = synth.camera_wfov()
    pose = synth.fill_frame_pose()
    Ri, ti = pose[:3, :3].T, -pose[:3, :3].T @ pose[:3, 3]
    res = {}
    for name, (y0, x0) in (("centre", (448, 448)), ("corner", (64, 704))):
        S = 128
        ys, xs = np.meshgrid(np.arange(y0, y0 + S), np.arange(x0, x0 + S), indexing="ij")
        xf, yf = xs.ravel().astype(np.float64), ys.ravel().astype(np.float64)
        dcam = np.stack([(xf - K[0, 2]) / K[0, 0], (yf - K[1, 2]) / K[1, 1], np.ones_like(xf)], 1)
        dcam /= np.linalg.norm(dcam, axis=1, keepdims=True)
        rays6 = np.concatenate([np.tile(ti, (len(xf), 1)), dcam @ Ri.T], 1)
        n = visits_per_ray(wide, rays6).reshape(S, S)
        row = {"mean_visits": float(n.mean()), "p99": float(np.percentile(n, 99)), "max": int(n.max())}
        for (th, tw) in ((1, 32), (2, 16), (4, 8), (8, 4), (16, 2), (32, 1)):
            t = n.reshape(S // th, th, S // tw, tw).transpose(0, 2, 1, 3).reshape(-1, th * tw)
            row["efficiency_%dx%d_rows_x_cols" % (th, tw)] = float(t.mean() / t.max(1).mean())
        srt = np.sort(n.ravel()).reshape(-1, 32)
        row["efficiency_sorted_by_cost_upper_bound"] = float(srt.mean() / srt.max(1).mean())
        res[name] = row
        print(name, json.dumps(row), flush=True)
    json.dump(res, open(os.path.join(ROOT, "profiles", "r1d_packet_study.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
