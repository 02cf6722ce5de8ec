#!/usr/bin/env python
"""Development diagnostics on a GPU box: every stage against the oracle, verbose."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj import Context, synth
from oracle import oracle as orc

def section(name):
    print(f"\n=== {name}", flush=True)

def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        print("!!! FAILED", fn.__name__, flush=True)

ctx = Context(0)
ctx.set_timing(True)

def t_radix():
    section("radix sort")
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 4096, 4097, 100003, 1 << 20):
        k = rng.integers(0, 1 << 30, n, dtype=np.uint32)
        if n > 1000: k[: n // 3] = k[0]      # duplicates
        v = np.arange(n, dtype=np.uint32)
        ks, vs = ctx.radix_sort(k, v)
        ref = np.argsort(k, kind="stable")
        print(n, "keys ok", np.array_equal(ks, k[ref]), "vals ok", np.array_equal(vs, v[ref]))

def t_compact():
    section("compaction")
    rng = np.random.default_rng(1)
    for shape in ((3, 4), (96, 128), (2, 97, 131), (720, 1280), (4, 720, 1280)):
        for dt in (np.float32, np.float64):
            h = rng.random(shape).astype(dt)
            thr = 0.5
            pix, I, counts = ctx.compact(h, thr)
            flat = h.reshape(-1)
            ref = np.nonzero(flat > (np.float32(thr) if dt == np.float32 else thr))[0]
            ok = np.array_equal(pix, ref.astype(np.uint32)) and np.array_equal(I, flat[ref].astype(np.float32))
            h3 = h.reshape((-1,) + h.shape[-2:])
            okc = np.array_equal(counts, (h3 > (np.float32(thr) if dt == np.float32 else thr)).reshape(len(h3), -1).sum(1))
            print(shape, dt.__name__, "n", len(pix), "ok", ok, "counts ok", okc)
    pix, I, c = ctx.compact(np.zeros((5, 7), np.float32), 0.5); print("empty", len(pix), c)
    pix, I, c = ctx.compact(np.ones((1024, 1024), np.float32), 0.5); print("dense", len(pix), np.array_equal(pix, np.arange(1 << 20, dtype=np.uint32)))

def check_cast(name, V, F, rays6, brute=True):
    ctx.set_mesh(V, F); ctx.build_bvh()
    st = ctx.stats()
    t, f = ctx.cast_rays(rays6)
    if brute:
        tr, fr = orc.cast_brute_f32(V, F, rays6)
    else:
        tr, fr = orc.Bvh(V, F).cast_f32(rays6)
    same_f = np.array_equal(f, fr)
    same_t = np.array_equal(t.view(np.uint32), tr.view(np.uint32))
    print(f"{name}: nF={len(F)} rays={len(rays6)} hits={int((fr>=0).sum())} nodes={st['n_wide_nodes']} depth={st['wide_depth']} build_ms={st['last_build_ms']:.3f} face_equal={same_f} t_equal={same_t}", flush=True)
    if not same_f:
        bad = np.nonzero(f != fr)[0]
        print("   mismatches:", len(bad), "first:", bad[:5], f[bad[:5]], fr[bad[:5]], t[bad[:5]], tr[bad[:5]])
    return same_f and same_t

def cam_rays(K, H, W, pose, step=1):
    ys, xs = np.mgrid[0:H:step, 0:W:step]
    xs = xs.reshape(-1).astype(np.int64); ys = ys.reshape(-1).astype(np.int64)
    xf = Context.frame_xform(K, pose)
    return orc.rays_object_frame(xs, ys, xf), xs, ys, xf

def t_cast():
    section("cast_rays vs brute force")
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    # single triangle known answer
    V = np.array([[-1, -1, 5], [1, -1, 5], [0, 1, 5]], np.float32); F = np.array([[0, 1, 2]], np.int32)
    r = np.array([[0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, -1], [5, 5, 0, 0, 0, 1]], np.float32)
    ctx.set_mesh(V, F); ctx.build_bvh(); print("single tri:", ctx.cast_rays(r))
    for name in ("tiny", "small", "c1_30k"):
        nu, nv = synth.MESH_CONFIGS[name]
        V, F = synth.param_mesh(nu, nv, seed=1)
        rays6, *_ = cam_rays(K, H, W, pose, step=8 if name != "c1_30k" else 6)
        check_cast(name, V, F, rays6, brute=True)
    # random incoherent rays
    rng = np.random.default_rng(5)
    V, F = synth.param_mesh(40, 25, seed=2)
    o = rng.normal(size=(20000, 3)).astype(np.float32) * 150
    d = -o + rng.normal(size=(20000, 3)).astype(np.float32) * 40
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    check_cast("random rays", V, F, np.hstack([o, d]).astype(np.float32))
    # degenerate: empty mesh, 1..5 triangles
    for n in (1, 2, 3, 4, 5, 9):
        Vs = rng.normal(size=(3 * n, 3)).astype(np.float32); Vs[:, 2] += 5
        Fs = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
        rr = np.zeros((500, 6), np.float32); rr[:, 3:] = rng.normal(size=(500, 3)) * 0.3; rr[:, 5] = 1
        check_cast(f"{n} tris", Vs, Fs, rr)
    ctx.set_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)); ctx.build_bvh()
    print("empty mesh:", ctx.cast_rays(np.array([[0, 0, 0, 0, 0, 1]], np.float32)))

def t_project():
    section("project (object + camera) vs oracle")
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    V, F = synth.param_mesh(150, 100, seed=0)
    heat = synth.gaussian_heatmap((H, W), dtype=np.float32)
    heat[:, :400] = synth.blob_heatmap((H, 400), seed=3)
    ctx.set_mesh(V, F); ctx.build_bvh()
    for thr in (0.5, 0.05):
        ctx.accum_reset()
        res = ctx.project(heat, K, pose[None], thr, "object", True, want=("pixel", "intensity", "t_hit", "face", "point", "point64"))
        xs, ys, I = orc.heatmap_to_points(heat, thr)
        xf = Context.frame_xform(K, pose)
        rays6 = orc.rays_object_frame(xs, ys, xf)
        bv = orc.Bvh(V, F)
        tr, fr = bv.cast_f32(rays6)
        hist, fmax, vmax = orc.accumulate(fr, I, F, len(V))
        gh, gf, gv = ctx.accum_get()
        dcam = orc.compute_rays(xs, ys, K)
        pref = dcam * tr.astype(np.float64)[:, None]
        hit = fr >= 0
        print(f"thr {thr}: n {res['n']} vs {len(xs)} hits {res['hits']} vs {int(hit.sum())}",
              "pix", np.array_equal(res["pixel"], (ys * W + xs).astype(np.uint32)),
              "I", np.array_equal(res["intensity"], I),
              "face", np.array_equal(res["face"], fr), "t", np.array_equal(res["t_hit"].view(np.uint32), tr.view(np.uint32)),
              "p64", np.array_equal(res["point64"][hit], pref[hit]),
              "p32", np.allclose(res["point"][hit], pref[hit], rtol=0, atol=1e-3),
              "hist", np.array_equal(gh, hist), "fmax", np.array_equal(gf, fmax), "vmax", np.array_equal(gv, vmax), flush=True)
        print("   timings", ctx.last_timings())
    # camera mode
    ctx.pose_mesh(pose)
    Vc = ctx.posed_vertices()
    Vref = orc.pose_vertices(V.astype(np.float64), pose)
    print("posed vertices equal:", np.array_equal(Vc, Vref))
    ctx.accum_reset()
    res = ctx.project(heat, K, None, 0.5, "camera", True, want=("pixel", "t_hit", "face", "point64"))
    xs, ys, I = orc.heatmap_to_points(heat, 0.5)
    dcam = orc.compute_rays(xs, ys, K)
    bvc = orc.Bvh(Vref, F)
    tr, fr = bvc.cast_f32(orc.rays6_camera(dcam))
    print("camera mode: face", np.array_equal(res["face"], fr), "t", np.array_equal(res["t_hit"].view(np.uint32), tr.view(np.uint32)),
          "hits", res["hits"], int((fr >= 0).sum()), "refit_ms", ctx.stats()["last_refit_ms"])

def t_perf():
    section("perf C2: 1024x1024 dense, 500k tris; and 1M tris")
    import torch
    K, H, W = synth.camera_wfov()
    pose = synth.fill_frame_pose()
    for name in ("c2_500k", "ns_1m"):
        nu, nv = synth.MESH_CONFIGS[name]
        V, F = synth.param_mesh(nu, nv, seed=0, scale=6.0)
        t0 = time.time(); ctx.set_mesh(V, F); ctx.build_bvh(); t1 = time.time()
        st = ctx.stats()
        print(name, "build wall", round(t1 - t0, 3), "device build_ms", st["last_build_ms"], "nodes", st["n_wide_nodes"], "depth", st["wide_depth"])
        heat = torch.ones((1, H, W), dtype=torch.float32, device="cuda")
        n = H * W
        out = dict(pixel=torch.empty(n, dtype=torch.int32, device="cuda"), t_hit=torch.empty(n, dtype=torch.float32, device="cuda"),
                   face=torch.empty(n, dtype=torch.int32, device="cuda"))
        ctx.set_stats(True)
        print("   n,hits", ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True))
        st = ctx.stats()
        print("   nodes/ray", st["nodes_fetched"] / max(1, st["rays"]), "tris/ray", st["tris_tested"] / max(1, st["rays"]))
        ctx.set_stats(False)
        for it in range(5):
            ctx.project_device(heat, K, pose[None], 0.5, "object", True, out=out, sync=True)
            tm = ctx.last_timings()
            print("   ", tm, "Mrays/s trace", n / tm["trace_ms"] / 1e3)
        # parity at full size vs oracle BVH
        xs = np.tile(np.arange(W, dtype=np.int64), H); ys = np.repeat(np.arange(H, dtype=np.int64), W)
        rays6 = orc.rays_object_frame(xs, ys, Context.frame_xform(K, pose))
        t0 = time.time(); bv = orc.Bvh(V, F); t1 = time.time(); tr, fr = bv.cast_f32(rays6); t2 = time.time()
        f = out["face"].cpu().numpy(); t = out["t_hit"].cpu().numpy()
        print("   oracle build s", round(t1 - t0, 2), "cast s", round(t2 - t1, 2), "face eq", np.array_equal(f, fr), "t eq",
              np.array_equal(t.view(np.uint32), tr.view(np.uint32)), "hit frac", (fr >= 0).mean(), flush=True)
        if not np.array_equal(f, fr):
            bad = np.nonzero(f != fr)[0]; print("   mismatches", len(bad), bad[:8], f[bad[:8]], fr[bad[:8]], t[bad[:8]], tr[bad[:8]])

which = sys.argv[1:] or ["radix", "compact", "cast", "project", "perf"]
for w in which:
    run(globals()["t_" + w])
print("\nDONE")
