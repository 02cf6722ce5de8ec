#!/usr/bin/env python
"""Headline benchmark: Mrays/s and ms/frame of the defect back-projection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh c2_500k]
                    [--scaling weak|strong] [--batches B] [--no-cpu] [--no-configs]

Workload (BASELINE.json configs[1]): Azure-Kinect WFOV 1024x1024 full-frame dense heatmap (1 048 576 rays
above the threshold) against a 500k-triangle mesh, BVH prebuilt, one frame per step:
threshold+compaction -> pixel rays (object frame) -> 8-wide BVH traversal -> histogram / max accumulation -> hit points.

* value      device-resident: heatmap already in HBM, outputs stay in HBM; per-step CUDA events on the
             launching stream, L2 flushed between steps (256 MiB memset, outside the events).
* e2e        the same frames through the host-buffer API (FrameStream over dp_project): pinned host heatmap in, the
             drop-in's payload out -- the float32 hit point and the face id of every ray (16 B/ray; the pixel list of
             a dense frame is the identity and is not shipped) + counts; copies inside the timed region.  Three more modes
             are reported beside it: `full` (+ t_hit, 20 B/ray), `lean` (t_hit + face, 8 B/ray) and `accumulate_only` (no
             per-ray read-back: the per-face histogram / maxima ARE the product for go.Mesh3d,
             /root/reference/src/web_vis.py:203-217).
* roofline   traversal kernel (k_trace): algorithmic bytes/ray (node bytes x nodes fetched + 48 B x triangles tested +
             32 B of ray I/O + 32 B of accumulator RMW per hit; counts measured live by the counting kernel variant;
             node bytes = 208 for the uncompressed node set traced while the hierarchy fits L2, 80 for the compressed
             one) / the kernel's mean duration (CUDA events), against MEASURED_PEAKS.json's HBM copy rate.
* configs    (N = 1) every other BASELINE.json config, driver-timed in the same run: north-star 1M triangles, configs[3]
             5M triangles (warm build + traversal), configs[2] 64 views (one launch / per-frame refit), configs[0].
* cpu_baseline / --impl reference   the reference's own CPU path: the real open3d RaycastingScene when importable
             (kind "reference"), else the CPU restatement (oracle/, OpenMP on every host core the process may use,
             kind "port") on the same frames: mesh posing + BVH build per call (as the reference does, :253-254) +
             rays + closest hit.
N > 1 (torchrun), --scaling weak (default): frames are sharded, every rank runs K frames against its own BVH replica.
The K frames form B batches (default 2); after each batch the accumulator block is snapshotted and combined over the
ranks (one SUM, one MAX all-reduce) and the batch's last frame's compacted hit records are gathered unpadded to rank 0
-- on a side stream, overlapped with the next batch's frames; only the last batch's combine is exposed.  The line also
carries `configs.c4_5m_strong` / `c2_500k_strong`: ONE frame's compacted ray list split over the ranks
(--scaling strong makes that the headline).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))

THR = 0.5
B_NODE_Q, B_NODE_FAT, B_TRI, B_RAY_IO, B_HIT_ACC = 80, 208, 48, 32, 32     # DESIGN.md "algorithmic bytes"
B_BUILD_TRI = 310                                                           # LBVH build, bytes per triangle
FAT_MAX_BYTES = 96 << 20                                                    # csrc/dp_internal.cuh


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_ncu_traffic(mesh):
    """dram bytes per launch of the traversal kernel from the committed ncu summary, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(mesh)
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons sampled with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.001)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload(mesh_name, scale=6.0):
    from defectproj import synth
    nu, nv = synth.MESH_CONFIGS[mesh_name]
    V, F = synth.param_mesh(nu, nv, seed=0, scale=scale)
    K, H, W = synth.camera_wfov()
    return V, F, K, H, W


def frame_pose(i):
    """Per-frame pose: the fill-frame camera nudged along the tube so successive frames differ."""
    from defectproj import synth
    a = 0.35 * np.sin(0.7 * i)
    return synth.look_at_pose(eye=(6 * 60.0 + 20 * a, -30.0 + 15 * np.cos(0.3 * i), 6 * 8.0),
                              target=(6 * 30.0, 6 * 52.0, 10.0 * a))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    """All the host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the
    OpenMP runtime of the CPU arm is told the real count explicitly."""
    from oracle import oracle as orc
    return orc.set_num_threads(len(os.sched_getaffinity(0)))


def open3d_or_none():
    try:
        import open3d as o3d                                        # the reference's own ray caster, when a wheel exists
        return o3d
    except Exception:
        return None


def cpu_frames(V, F, K, H, W, budget_s, max_frames, first_frame=0):
    """The reference path on the CPU, all host threads.  Returns per-frame seconds and ray counts."""
    from oracle import oracle as orc
    cores = cpu_threads()
    o3d = open3d_or_none()
    heat = np.ones((H, W), np.float32)
    V64 = V.astype(np.float64)
    t_total, rays, frames, t_cast = 0.0, 0, 0, 0.0
    t_start = time.perf_counter()
    while frames < max_frames and (frames == 0 or time.perf_counter() - t_start < budget_s):
        pose = frame_pose(first_frame + frames)
        t0 = time.perf_counter()
        if o3d is not None:
            # the reference's own lines: posed legacy mesh -> from_legacy -> RaycastingScene (rebuilt per call) -> cast_rays
            xs, ys, _ = orc.heatmap_to_points(heat, THR)
            d = orc.compute_rays(xs, ys, K)
            Vp = V64 @ pose[:3, :3].T + pose[:3, 3]
            legacy = o3d.geometry.TriangleMesh(o3d.utility.Vector3dVector(Vp), o3d.utility.Vector3iVector(F))
            mesh = o3d.t.geometry.TriangleMesh.from_legacy(legacy)
            rays_t = o3d.core.Tensor(np.hstack((np.zeros_like(d), d)), dtype=o3d.core.Dtype.Float32)
            scene = o3d.t.geometry.RaycastingScene()
            scene.add_triangles(mesh)
            t1 = time.perf_counter()
            scene.cast_rays(rays_t)
            n = len(xs)
        else:
            Vc = orc.pose_vertices(V64, pose)               # :549-550 + :245
            bvh = orc.Bvh(Vc, F)                            # :253-254, rebuilt every call like the reference
            t1 = time.perf_counter()
            n = bvh.project_frame(heat, THR, K)["n"]        # :551-556
            del bvh
        t2 = time.perf_counter()
        t_total += t2 - t0
        t_cast += t2 - t1
        rays += n
        frames += 1
    return dict(seconds=t_total, cast_seconds=t_cast, rays=rays, frames=frames, cores=cores,
                kind="reference" if o3d is not None else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    V, F, K, H, W = workload(args.mesh)
    # bounded: every step is one full frame unless that would run for minutes; then a row band of it
    probe = cpu_frames(V, F, K, H, W, 0.0, 1)
    per = probe["seconds"]
    rows = H
    if per * (args.steps + args.warmup) > 150.0:
        rows = max(16, int(H * 150.0 / (per * (args.steps + args.warmup))))
    Hs = rows
    for _ in range(args.warmup):
        cpu_frames(V, F, K, Hs, W, 0.0, 1)
    t0 = time.perf_counter()
    tot = cpu_frames(V, F, K, Hs, W, 1e9, args.steps, first_frame=1)
    el = time.perf_counter() - t0
    val = tot["rays"] / el / 1e6
    who = ("open3d RaycastingScene (Embree), the reference's own path" if tot["kind"] == "reference"
           else "CPU restatement of the reference path (oracle/, OpenMP)")
    sample = (f"{who}; {args.steps} frames of {Hs}x{W} dense rays ({'full frame' if Hs == H else 'top row band of the 1024x1024 frame'})"
              f" vs {len(F)} triangles; per frame: float64 vertex posing + BVH build (rebuilt per call, as the reference does) "
              f"+ threshold + rays + closest hit + accumulation; cast-only {tot['rays'] / tot['cast_seconds'] / 1e6:.2f} Mrays/s; "
              f"{tot['cores']} threads (sched_getaffinity: {len(os.sched_getaffinity(0))}, OMP_NUM_THREADS in the environment: "
              f"{os.environ.get('OMP_NUM_THREADS', 'unset')})")
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, len(F), H, W),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": tot["cores"], "kind": tot["kind"], "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args, nF, H, W):
    return {"workload": f"configs[1]: Azure Kinect WFOV {H}x{W} full-frame dense projection ({H * W} rays/frame) on a "
                        f"{nF}-triangle mesh ({args.mesh}), threshold {THR}, BVH prebuilt, 1 frame per step",
            "mesh": args.mesh, "triangles": nF, "rays_per_frame": H * W,
            "l2": "flushed between timed steps by a 256 MiB memset outside the event pairs",
            "frame": "object (rays through the inverse pose, static BVH)"}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on (and first-touch the pinned buffers from) the CPUs of the GPU's own NUMA node,
    so that eight ranks do not all stream their frames through one socket's memory.  Returns a short description
    for the bench line; does nothing when the topology is not visible or the node has none of our CPUs."""
    if os.environ.get("DP_NUMA_BIND", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        mine = cpus & os.sched_getaffinity(0)
        if node < 0 or not mine or mine == os.sched_getaffinity(0):
            return f"node {node}: no narrower cpu set"
        os.sched_setaffinity(0, mine)
        return f"node {node}: {len(mine)} cpus"
    except Exception as e:                                  # topology not visible in this container
        return f"unavailable ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------ helpers (GPU arm)
def node_bytes(st):
    """Bytes of one node of the set the traversal actually reads (csrc/api.cu view_of)."""
    fat = st["n_wide_nodes"] * B_NODE_FAT + st["n_tris"] * B_TRI <= FAT_MAX_BYTES and os.environ.get("DP_FAT", "1") != "0"
    return B_NODE_FAT if fat else B_NODE_Q


def bytes_per_ray(st_counts, st):
    n = max(1, st_counts["rays"])
    nodes, tris, hit = st_counts["nodes_fetched"] / n, st_counts["tris_tested"] / n, st_counts["hits"] / n
    nb = node_bytes(st)
    return nb * nodes + B_TRI * tris + B_RAY_IO + B_HIT_ACC * hit, nodes, tris, hit, nb


def timed_traversal(ctx, heat, K, poses, out, flush, steps, warmup=3):
    """(ms per frame, ms of k_trace) over `steps` frames, L2 flushed before each, CUDA events on the current stream."""
    import torch
    stream = torch.cuda.current_stream()
    for i in range(warmup):
        flush.zero_()
        ctx.project_device(heat, K, poses[i % len(poses)][None], THR, "object", True, out=out, sync=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    k_ms = []
    for i in range(steps):
        flush.zero_()
        sampled = i % 4 == 3
        if sampled:
            ctx.set_timing(True)
        ev[i][0].record(stream)
        ctx.project_device(heat, K, poses[(warmup + i) % len(poses)][None], THR, "object", True, out=out, sync=False)
        ev[i][1].record(stream)
        if sampled:
            k_ms.append(ctx.last_timings()["trace_ms"])
            ctx.set_timing(False)
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev])), float(np.mean(k_ms))


def sub_config_dense(mesh, peak, flush, steps=12):
    """One other BASELINE config on this GPU: warm build (median of 5) + dense-frame traversal + its roofline."""
    import torch
    from defectproj import Context
    V, F, K, H, W = workload(mesh)
    n_pix = H * W
    with Context(torch.cuda.current_device()) as ctx:
        ctx.set_mesh(V, F).build_bvh()
        builds = []
        for _ in range(5):
            ctx.build_bvh()
            builds.append(ctx.stats()["last_build_ms"])
        heat = torch.ones((1, H, W), dtype=torch.float32, device="cuda")
        out = dict(pixel=torch.empty(n_pix, dtype=torch.int32, device="cuda"), intensity=torch.empty(n_pix, device="cuda"),
                   t_hit=torch.empty(n_pix, device="cuda"), face=torch.empty(n_pix, dtype=torch.int32, device="cuda"),
                   point=torch.empty((n_pix, 3), device="cuda"))
        poses = [frame_pose(i) for i in range(steps + 3)]
        ctx.set_stats(True)
        ctx.accum_reset()
        n_rays, n_hits = ctx.project_device(heat, K, poses[0][None], THR, "object", True, out=out, sync=True)
        st = ctx.stats()
        ctx.set_stats(False)
        b_ray, nodes, tris, hit, nb = bytes_per_ray(st, st)
        ms, k_ms = timed_traversal(ctx, heat, K, poses, out, flush, steps)
        build_ms = float(np.median(builds))
        achieved = n_rays * b_ray / (k_ms * 1e-3) / 1e9
        b_gbs = B_BUILD_TRI * len(F) / (build_ms * 1e-3) / 1e9
        return {"triangles": len(F), "wide_nodes": st["n_wide_nodes"], "bvh_bytes": st["n_wide_nodes"] * nb + st["n_tris"] * B_TRI,
                "node_bytes": nb, "build_ms_warm": build_ms, "build_ms_all": [round(b, 4) for b in builds],
                "build_mtris_s": len(F) / build_ms / 1e3, "build_gbs": b_gbs, "build_frac": b_gbs / peak,
                "ms_per_frame": ms, "mrays_s": n_rays / ms / 1e3, "trace_ms": k_ms, "trace_mrays_s": n_rays / k_ms / 1e3,
                "nodes_per_ray": nodes, "tris_per_ray": tris, "hit_frac": hit, "bytes_per_ray": b_ray,
                "roofline_frac": achieved / peak, "rays_per_frame": n_rays}


def sub_config_c3(flush):
    """configs[2]: 64 Fibonacci views x 720p blob heatmaps, 500k triangles, histogram accumulated over the views."""
    import torch
    from defectproj import Context, synth
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c2_500k"], seed=0)
    K, H, W = synth.camera_720p()
    B = 64
    poses = synth.fibonacci_poses(B, radius=600.0)
    heats = torch.from_numpy(np.stack([synth.blob_heatmap((H, W), seed=200 + i) for i in range(B)])).cuda()
    stream = torch.cuda.current_stream()
    with Context(torch.cuda.current_device()) as ctx:
        ctx.set_mesh(V.astype(np.float64), F).build_bvh()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = {}
        for rep in range(3):                                   # one launch for the 64 views (object frame)
            ctx.accum_reset()
            flush.zero_()
            e0.record(stream)
            n, h = ctx.project_device(heats, K, poses, THR, "object", True, sync=True)
            e1.record(stream)
            torch.cuda.synchronize()
            res["object_one_launch_ms"] = e0.elapsed_time(e1)
        res["rays"], res["hits"] = n, h
        hist_obj = ctx.accum_get()[0]
        for rep in range(2):                                   # reference-literal: per-frame float64 posing + refit + launch
            ctx.accum_reset()
            flush.zero_()
            e0.record(stream)
            for b in range(B):
                ctx.pose_mesh(poses[b], stream)
                ctx.project_device(heats[b:b + 1], K, None, THR, "camera", True, sync=False)
            e1.record(stream)
            torch.cuda.synchronize()
            res["camera_refit_ms_per_frame"] = e0.elapsed_time(e1) / B
        res["refit_ms"] = ctx.stats()["last_refit_ms"]
        hist_cam = ctx.accum_get()[0]
        res["object_ms_per_frame"] = res["object_one_launch_ms"] / B
        res["hist_faces_differing_object_vs_camera"] = int((hist_obj != hist_cam).sum())
        res["note"] = "tests/test_gpu_northstar.py::test_config3_64_views_refit_vs_object_frame asserts the differing rays are ties"
        return res


def sub_config_c1(flush):
    """configs[0]: 30k-triangle mesh, one 720p heatmap (Gaussian: 10 885 rays; dense: 921 600), fixed pose."""
    import torch
    from defectproj import Context, synth
    V, F = synth.param_mesh(*synth.MESH_CONFIGS["c1_30k"], seed=0)
    K, H, W = synth.camera_720p()
    pose = synth.fixed_pose()
    stream = torch.cuda.current_stream()
    res = {}
    with Context(torch.cuda.current_device()) as ctx:
        ctx.set_mesh(V, F).build_bvh()
        for name, heat in (("gaussian", synth.gaussian_heatmap((H, W), dtype=np.float32)), ("dense", np.ones((H, W), np.float32))):
            hd = torch.from_numpy(heat).cuda()[None]
            ts = []
            for rep in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.project_device(hd, K, pose[None], THR, "object", True, sync=False)
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            n, h = ctx.project_device(hd, K, pose[None], THR, "object", True, sync=True)
            res[name] = {"rays": n, "hits": h, "ms_per_frame": float(np.median(ts[2:]))}
    return res


def strong_scaling(mesh, rank, world, local, steps, flush):
    """ONE dense frame per step, its compacted ray list split over the ranks (Projector.project_frame_sharded):
    replicated compaction, sharded ray generation / traversal / accumulation / hit points, then the slices of
    t_hit and face travel to every rank (one broadcast per rank and array) and the accumulator snapshot is combined."""
    import torch
    import torch.distributed as dist
    from defectproj import Projector
    V, F, K, H, W = workload(mesh)
    n_pix = H * W
    dev = f"cuda:{local}"
    proj = Projector(V, F, device=local)
    stream = torch.cuda.current_stream()
    heat = torch.ones((H, W), dtype=torch.float32, device=dev)
    out = dict(t_hit=torch.empty(n_pix, device=dev), face=torch.empty(n_pix, dtype=torch.int32, device=dev))
    poses = [frame_pose(i) for i in range(steps + 3)]            # the SAME frame on every rank
    # peer-mapped result windows: the traversal stores its slice into every rank's arrays, no collective per frame
    peer = world > 1 and proj.peer is not None and proj.enable_peer(result_rays=n_pix)
    peer_res = None
    if peer:
        pout = {}
        for i in range(3):
            proj.project_frame_sharded(heat, K, poses[i], THR, out=pout, gather="peer", reduce=False, reset=False)
        torch.cuda.synchronize()
        dist.barrier()
        pms = []
        for i in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            pn, ph, _ = proj.project_frame_sharded(heat, K, poses[3 + i], THR, out=pout, gather="peer", reduce=False, reset=False)
            e1.record(stream)
            torch.cuda.synchronize()
            pms.append(e0.elapsed_time(e1))
        # the same frames as a sequence: nothing waits between frames (slots alternate), one event pair around all of them
        dist.barrier()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for i in range(steps):
            proj.project_frame_sharded(heat, K, poses[3 + i], THR, out=pout, gather="peer", reduce=False, reset=False)
        p1.record(stream)
        torch.cuda.synchronize()
        seq_ms = p0.elapsed_time(p1) / steps
        ref_face = torch.empty(n_pix, dtype=torch.int32, device=dev)       # the last frame, unsharded, on this rank
        proj.ctx.set_ray_shard(0, 1)
        proj.ctx.project_device(heat[None], K, poses[3 + steps - 1][None], THR, "object", False, out=dict(face=ref_face), sync=True)
        same = bool(torch.equal(ref_face, pout["face"][:n_pix])) and proj.ctx.peer_status() == 0
        tt = torch.tensor([float(np.mean(pms)), seq_ms, 0.0 if same else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        peer_res = {"ms_per_frame": float(tt[0]), "mrays_s": n_pix / float(tt[0]) / 1e3, "sequence_ms_per_frame": float(tt[1]),
                    "sequence_mrays_s": n_pix / float(tt[1]) / 1e3, "whole_frame_on_every_rank_equals_unsharded": float(tt[2]) == 0.0,
                    "how": "csrc/peer.cu: k_trace stores its slice of t_hit / face into the result window of EVERY rank (NVLink "
                           "stores from the traversal's epilogue), then a one-warp flag barrier; no collective, no host wait"}
        proj.ctx.accum_reset(stream)
        torch.cuda.synchronize()
        dist.barrier()
    for i in range(3):
        proj.project_frame_sharded(heat, K, poses[i], THR, out=out, reduce=False, reset=False)
    proj.combiner.submit()
    proj.combined()
    proj.combiner.reset_totals()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the same frames as a SEQUENCE: frame i's gathers run on a side stream under frame i+1's kernels (double-buffered
    # results), nothing waits per frame; only for hierarchies larger than L2 (no flush needed between frames)
    pipelined_ms = None
    if len(F) * B_TRI > (126 << 20):
        outs = [out, dict(t_hit=torch.empty(n_pix, device=dev), face=torch.empty(n_pix, dtype=torch.int32, device=dev))]
        side = proj.combiner.side
        for rep in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            for i in range(steps):
                if i >= 2:
                    stream.wait_stream(side)          # the buffer about to be overwritten has been gathered (frame i-2)
                proj.project_frame_sharded(heat, K, poses[3 + i], THR, out=outs[i & 1], reduce=False, reset=False,
                                           gather_stream=side)
            stream.wait_stream(side)
            p1.record(stream)
            torch.cuda.synchronize()
            pipelined_ms = p0.elapsed_time(p1) / steps
        proj.ctx.accum_reset(stream)
        torch.cuda.synchronize()
    ms = []
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hits_local = 0
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n, h, rng = proj.project_frame_sharded(heat, K, poses[3 + i], THR, out=out, reduce=False, reset=False)
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        hits_local += int(h)
    # the accumulators are combined ONCE for the batch of frames (not per frame)
    e_all0.record(stream)
    proj.combiner.submit()
    hist = proj.combined()[0]
    e_all1.record(stream)
    torch.cuda.synchronize()
    combine_ms = e_all0.elapsed_time(e_all1)
    t = torch.tensor([float(np.mean(ms)) + combine_ms / steps, combine_ms, pipelined_ms or 0.0], dtype=torch.float64, device=dev)
    hl = torch.tensor([hits_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(hl)
    pm = float(t[2]) if pipelined_ms else None
    h = int(hl[0]) // steps
    ok = int(hist.sum().item()) == int(hl[0]) and h == int((out["face"][:n] >= 0).sum().item())
    proj.close()
    return {"triangles": len(F), "rays_per_frame": n, "hits": h, "ms_per_frame": float(t[0]), "mrays_s": n / float(t[0]) / 1e3,
            "peer": peer_res,
            "combine_ms_per_batch": float(t[1]), "frames": steps,
            "pipelined_ms_per_frame": pm, "pipelined_mrays_s": (n / pm / 1e3) if pm else None,
            "slots_of_rank0": list(rng) if rank == 0 else None, "hist_total_equals_hits_equals_gathered_faces": bool(ok),
            "timed": "blocking call per frame (the ray count is read back), incl. ONE all-gather per result array (t_hit, "
                     "face) that leaves every rank with the whole frame; the accumulator block is combined once for the "
                     "batch of frames (its time / frames is added); L2 flushed before each frame outside the events; max "
                     "over ranks.  pipelined_*: the same frames as a sequence, frame i's gathers on a side stream under frame "
                     "i+1's kernels, one event pair around all frames (hierarchies larger than L2 only: no flush needed)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single process"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(local)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from defectproj import FrameStream, Projector
    from defectproj.projector import gather_hits

    V, F, K, H, W = workload(args.mesh)
    n_pix = H * W
    proj = Projector(V, F, device=local)
    ctx = proj.ctx
    comb = proj.combiner
    stream = torch.cuda.current_stream()
    dev = f"cuda:{local}"
    heat = torch.ones((1, H, W), dtype=torch.float32, device=dev)
    out = dict(pixel=torch.empty(n_pix, dtype=torch.int32, device=dev), intensity=torch.empty(n_pix, device=dev),
               t_hit=torch.empty(n_pix, device=dev), face=torch.empty(n_pix, dtype=torch.int32, device=dev),
               point=torch.empty((n_pix, 3), device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # every rank walks its own frames of the sweep
    poses = [frame_pose(rank * (args.steps + args.warmup) + i) for i in range(args.steps + args.warmup)]

    # -- per-ray node / triangle counts for the roofline (counting kernel variant, untimed)
    ctx.set_stats(True)
    ctx.accum_reset(stream)
    n_rays, n_hits = ctx.project_device(heat, K, poses[0][None], THR, "object", True, out=out, sync=True)
    st = ctx.stats()
    ctx.set_stats(False)
    b_ray, nodes_per_ray, tris_per_ray, hit_frac, nb = bytes_per_ray(st, st)

    def step_device(i):
        ctx.project_device(heat, K, poses[i][None], THR, "object", True, out=out, sync=False)

    # the K steps form `nb_batches` batches; each ends with a combine (accumulator snapshot + last frame's hit records)
    nb_batches = max(1, min(args.batches, args.steps))
    bounds = [round(b * args.steps / nb_batches) for b in range(nb_batches + 1)]
    rec_buf = [torch.empty((n_pix, 3), dtype=torch.int32, device=dev) for _ in range(2)]
    rec_cnt = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)]
    ev_pack = [torch.cuda.Event() for _ in range(2)]
    gathered = [0]
    # peer-mapped exchange windows (csrc/peer.cu): the per-batch combine is ONE kernel of the library per rank
    peer = world > 1 and proj.peer is not None and proj.enable_peer(record_rows=n_pix)
    comb = proj.combiner
    peer_gathered = torch.empty((world * n_pix, 3), dtype=torch.int32, device=dev) if peer and rank == 0 else None
    peer_rows = torch.zeros(1, dtype=torch.int64).pin_memory()

    def batch_end(b):
        """compute stream: hit records of the batch's last frame + accumulator snapshot; the reductions go to the side stream"""
        if peer:
            k = comb.acquire(stream)
            rec, cnt = comb.records(k)
            ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n_rays, out=rec, count_async=cnt,
                                    sync=False, stream=stream)
            comb.submit(stream, gather_root=0, gathered=peer_gathered, count_async=peer_rows)
            return
        k = b & 1
        ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n_rays, out=rec_buf[k], count_async=rec_cnt[k],
                                sync=False, stream=stream)
        ev_pack[k].record(stream)
        comb.submit(stream)

    def batch_gather(b):
        """side stream: the batch's hit records to rank 0, unpadded (the host reads the record count first)"""
        if peer:
            return                           # the records were pulled by rank 0's combine kernel
        k = b & 1
        ev_pack[k].synchronize()
        m = int(rec_cnt[k][0])
        with torch.cuda.stream(comb.side):
            comb.side.wait_event(ev_pack[k])
            got = gather_hits(rec_buf[k][:m], dst=0)
            if got is not None:
                gathered[0] = int(got.shape[0])

    for i in range(args.warmup):
        flush.zero_()
        step_device(i)
    # warm-up of the per-batch combine too (the first NCCL call of each kind sets up its channels, the first use of a
    # torch kernel loads its module)
    for b in range(2):
        batch_end(b)
        batch_gather(b)
    comb.result()
    ctx.accum_reset(stream)
    comb.reset_totals()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_tail = torch.cuda.Event(enable_timing=True)
    trace_ms = []
    for b in range(nb_batches):
        for i in range(bounds[b], bounds[b + 1]):
            if b > 0 and i == min(bounds[b] + 3, bounds[b + 1] - 1):
                batch_gather(b - 1)          # queued behind batch b-1's reductions on the side stream: overlaps this batch
            flush.zero_()
            sampled = i % 16 == 15 or i == args.steps - 1
            if sampled:
                ctx.set_timing(True)         # stage events inside the library on this step only (14 us per call)
            ev[i][0].record(stream)
            step_device(args.warmup + i)
            last_of_batch = i == bounds[b + 1] - 1
            if last_of_batch:
                batch_end(b)                 # inside the step's event pair: pack + vertex maxima + snapshot
            ev[i][1].record(stream)
            if sampled:
                # kernel-only duration of the traversal launch of this step
                trace_ms.append(ctx.last_timings()["trace_ms"])
                ctx.set_timing(False)
    batch_gather(nb_batches - 1)
    stream.wait_stream(comb.side)            # the last batch's reductions and gather are exposed
    ev_tail.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks_value = dict(sampler.stop())             # the 20 timed steps last ~6 ms: one or two NVML samples
    sampler = ClockSampler(local)                   # ... so the end-to-end runs (timed regions too) are sampled as well
    sampler.start()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tail_ms = ev[-1][1].elapsed_time(ev_tail)
    total_ms = float(sum(step_ms)) + tail_ms
    frames_ms = float(sum(step_ms))              # this rank's own frames; the spread over the ranks is what the tail waits for
    if peer:
        comb.check()
        gathered[0] = int(peer_rows[0])
    hist_t, fmax_t, vmax_t = comb.result()
    hist_total = int(hist_t.sum().item())                         # all ranks, all batches
    hits_sum = torch.tensor([0], dtype=torch.int64, device=dev)

    # -- the same combine, serial on the compute stream (nothing overlapped), for the record
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    if world > 1:
        dist.barrier()
    for rep in range(2):                      # twice: the first pass allocates the gather buffer on this stream's pool
        if world > 1:
            dist.barrier()
        e0.record(stream)
        if peer:
            k = comb.acquire(stream)
            rec, cnt = comb.records(k)
            ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n_rays, out=rec, count_async=cnt, sync=False, stream=stream)
            comb.submit(stream, reset=False, gather_root=0, gathered=peer_gathered, count_async=peer_rows)
            stream.wait_stream(comb.side)
            e1.record(stream)
        else:
            comb.submit(stream, reset=False)
            stream.wait_stream(comb.side)
            e1.record(stream)
            rec = ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n_rays, out=rec_buf[0], stream=stream)
            gather_hits(rec, dst=0)
        e2.record(stream)
        torch.cuda.synchronize()
        serial_reduce_ms, serial_gather_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    nccl_reduce_ms = nccl_gather_ms = None
    if peer:
        # the same combine through NCCL (two all-reduces; count exchange + all-gather of the records), for comparison
        from defectproj.projector import BatchCombiner
        nccl = BatchCombiner(ctx, None)
        for rep in range(3):
            dist.barrier()
            e0.record(stream)
            nccl.submit(stream, reset=False)
            stream.wait_stream(nccl.side)
            e1.record(stream)
            rec = ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n_rays, out=rec_buf[0], stream=stream)
            gather_hits(rec, dst=0)
            e2.record(stream)
            torch.cuda.synchronize()
            nccl_reduce_ms, nccl_gather_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)

    # -- end to end through the host-buffer API (defectproj.FrameStream -> dp_project): every frame's heatmap comes from
    #    pinned host memory and its per-ray results + counts go back to pinned host memory; H2D(i+1) | kernels(i) |
    #    D2H(i-1) overlap on three streams.  L2 is flushed on the kernel stream before every frame, INSIDE the timed region.
    e2e_steps = max(8, min(args.steps, 100))
    h_heats = [torch.ones((H, W), dtype=torch.float32).pin_memory() for _ in range(4)]
    e_poses = np.stack([poses[args.warmup + (i % args.steps)] for i in range(e2e_steps)])

    def flush_l2(_stream):
        flush[:132 << 20].zero_()

    def run_mode(want):
        fs = FrameStream(ctx, H, W, want=want)

        def run_stream(nf):
            rays = 0
            for i, res in fs.run([h_heats[i % 4] for i in range(nf)], K, e_poses[:nf], THR, "object", True, before_kernels=flush_l2):
                rays += res["n"]
            return rays
        run_stream(4)                                   # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rays = run_stream(e2e_steps)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        return dict(rays=rays, ms=float(fs.last_elapsed_ms), wall=wall, d2h=fs.last_d2h_bytes)

    # payload: what north_star names as the per-ray outputs -- hit point and face id (t_hit is |point|, not shipped);
    # full: t_hit as well (20 B/ray); lean: t_hit + face only; accumulate_only: nothing per ray
    modes = {"payload": ("pixel", "face", "point"), "full": ("pixel", "t_hit", "face", "point"),
             "lean": ("pixel", "t_hit", "face"), "accumulate_only": ()}
    e2e = {k: run_mode(w) for k, w in modes.items()}
    clocks_e2e = sampler.stop()
    clocks = {"sm_mhz": clocks_value["sm_mhz"] if clocks_value["sm_mhz"] is not None else clocks_e2e["sm_mhz"],
              "sm_max_mhz": clocks_value["sm_max_mhz"], "reasons": sorted(set(clocks_value["reasons"]) | set(clocks_e2e["reasons"])),
              "samples": clocks_value["samples"], "e2e_sm_mhz": clocks_e2e["sm_mhz"], "e2e_samples": clocks_e2e["samples"]}

    # what the link gives: the payload's bytes as plain pinned D2H copies back to back (no kernels, nothing else on the bus)
    pay_d = torch.empty(n_pix * 4, dtype=torch.float32, device=dev)
    pay_h = torch.empty(n_pix * 4, dtype=torch.float32).pin_memory()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        c0.record(stream)
        for _ in range(8):
            pay_h.copy_(pay_d, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize()
    pcie_d2h_gbs = 8 * pay_h.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del pay_d, pay_h
    h2d = n_pix * 4 + 128                       # heatmap + per-frame constants
    # the same frame as one blocking call (no overlap), for reference
    h_out = {"pixel": torch.empty(n_pix, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
             "t_hit": torch.empty(n_pix, dtype=torch.float32).pin_memory().numpy(),
             "face": torch.empty(n_pix, dtype=torch.int32).pin_memory().numpy(),
             "point": torch.empty((n_pix, 3), dtype=torch.float32).pin_memory().numpy()}
    heat_np = h_heats[0].numpy()[None]
    for i in range(3):
        ctx.project(heat_np, K, poses[i][None], THR, "object", True, out=h_out)
    t0 = time.perf_counter()
    for i in range(20):
        ctx.project(heat_np, K, poses[args.warmup + (i % args.steps)][None], THR, "object", True, out=h_out)
    blocking_ms = 1e3 * (time.perf_counter() - t0) / 20

    # -- max over ranks
    keys = list(modes)
    if world > 1:
        t = torch.tensor([total_ms, tail_ms, serial_reduce_ms, serial_gather_ms, frames_ms, -frames_ms, nccl_reduce_ms or 0.0,
                          nccl_gather_ms or 0.0] + [e2e[k]["ms"] for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, tail_ms, serial_reduce_ms, serial_gather_ms = (float(x) for x in t[:4])
        skew_ms = float(t[4]) + float(t[5])          # slowest rank's frames - fastest rank's frames
        if peer:
            nccl_reduce_ms, nccl_gather_ms = float(t[6]), float(t[7])
        for j, k in enumerate(keys):
            e2e[k]["ms"] = float(t[8 + j])
        c = torch.tensor([n_rays * args.steps, n_hits * 0] + [e2e[k]["rays"] for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays_total = float(c[0])
        for j, k in enumerate(keys):
            e2e[k]["rays_total"] = float(c[2 + j])
        # hits of every frame of every rank (the dense fill-frame: one count per step, read from the library)
        hits_sum[0] = int(hist_total)
    else:
        skew_ms = 0.0
        rays_total = float(n_rays * args.steps)
        for k in keys:
            e2e[k]["rays_total"] = float(e2e[k]["rays"])
    # the all-reduced histogram must count every hit of every step of every rank once: compare with the hit counts
    # the kernels reported (dp_project's n_hits), summed over steps and ranks
    hits_reported = torch.tensor([0], dtype=torch.int64, device=dev)
    ctx.accum_reset(stream)
    chk = 0
    for i in range(args.steps):
        chk += ctx.project_device(heat, K, poses[args.warmup + i][None], THR, "object", False, out=out, sync=True)[1]
    hits_reported[0] = chk
    if world > 1:
        dist.all_reduce(hits_reported)

    sub = {}
    if world > 1 and not args.no_configs:
        for mesh in ("c4_5m", "c2_500k"):
            sub[mesh + "_strong"] = strong_scaling(mesh, rank, world, local, 10, flush)

    if rank == 0:
        peak, peak_src = load_peaks()
        k_ms = float(np.mean(trace_ms))
        achieved = n_rays * b_ray / (k_ms * 1e-3) / 1e9
        achieved80 = n_rays * (b_ray - (nb - B_NODE_Q) * nodes_per_ray) / (k_ms * 1e-3) / 1e9
        st2 = ctx.stats()

        def e2e_entry(k):
            m = e2e[k]
            return {"value": m["rays_total"] / (m["ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": m["ms"] / e2e_steps,
                    "d2h_bytes_per_step": m["d2h"], "wall_ms_per_frame": 1e3 * m["wall"] / e2e_steps, "outputs": list(modes[k])}
        head = e2e_entry("payload")
        line = {
            "metric": "Mrays/s", "value": rays_total / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, len(F), H, W),
            "clocks": clocks,
            "numa": numa,
            "e2e": {"value": head["value"], "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": head["d2h_bytes_per_step"],
                    "ms_per_frame": head["ms_per_frame"], "steps": e2e_steps, "wall_ms_per_frame": head["wall_ms_per_frame"],
                    "blocking_call_ms_per_frame": blocking_ms,
                    "pcie_d2h_gbs": pcie_d2h_gbs,
                    "pcie_bound_mrays_s": pcie_d2h_gbs * 1e9 / 16.0 / 1e6,
                    "api": "defectproj.FrameStream.run (3-stream pipeline over dp_project); blocking_call = Context.project",
                    "l2": "132 MiB (> 126 MB L2) memset on the kernel stream before every frame, inside the timed region",
                    "pcie": "pcie_d2h_gbs: this rank's pinned device-to-host copy rate measured with the payload's bytes back to back; "
                            "pcie_bound_mrays_s = that rate / 16 B per ray: the ceiling of the payload mode on this link",
                    "outputs": "the drop-in's payload: the float32 hit point and the face id of every ray (16 B/ray; the reference "
                               "returns the hit points, /root/reference/src/defect_projection.py:261-264, the face ids are the "
                               "extension north_star names) + ray/hit counts; t_hit (= |point|) travels in mode 'full' "
                               "(20 B/ray); pixel u32 is the identity for a dense frame (synthesised on the host, not copied); "
                               "heatmap f32 in; pinned host memory",
                    "modes": {k: e2e_entry(k) for k in keys}},
            # k_compact_project (the compaction, which also does the call's resets and uploads) and k_trace per frame (rays
            # and hit points are generated inside k_trace; DP_FUSE_RAYS=0: also k_raygen and k_points; DP_FUSED_PROLOGUE=0:
            # also k_project_prologue); per batch k_pack_records, k_vertex_max, and k_peer_snapshot +
            # k_peer_combine (peer path) or a snapshot copy followed by NCCL's kernels
            "gpu_launches": ((2 if os.environ.get("DP_FUSED_PROLOGUE", "1") != "0" else 3)
                             + (0 if os.environ.get("DP_FUSE_RAYS", "1") != "0" else 2)) * args.steps + (4 if peer else 3) * nb_batches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_ncu_traffic(args.mesh), "peak_source": peak_src,
                         "kernel": "k_trace<false,0,%d,%d>" % ((6, 1) if nb == B_NODE_FAT else (7, 0)), "kernel_ms": k_ms,
                         "bytes_per_ray": b_ray, "node_bytes": nb,
                         "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray, "hit_frac": hit_frac,
                         "frac_at_80B_nodes": achieved80 / peak,
                         "note": "the traversal reads the UNCOMPRESSED 208-byte node set while nodes + records fit L2 (80-byte "
                                 "compressed nodes beyond): bytes_per_ray uses the struct actually read, frac_at_80B_nodes "
                                 "restates the fraction with round 1's 80-byte constant for comparison.  The set is smaller "
                                 "than L2, so most fetched bytes are L2/L1 hits: frac is algorithmic bytes over the HBM copy "
                                 "rate (it may exceed 1), not DRAM traffic"},
            "ms_per_frame": total_ms / args.steps,
            "bvh": {"build_ms": st2["last_build_ms"], "wide_nodes": st2["n_wide_nodes"], "depth": st2["wide_depth"],
                    "bytes": st2["n_wide_nodes"] * nb + st2["n_tris"] * B_TRI},
            "combine": {"batches": nb_batches, "path": "peer memory (csrc/peer.cu)" if peer else ("nccl" if world > 1 else "local"),
                        "exposed_tail_ms": tail_ms, "rank_spread_of_frames_ms": skew_ms,
                        "serial_reduce_ms": serial_reduce_ms, "serial_gather_ms": serial_gather_ms,
                        "nccl_serial_reduce_ms": nccl_reduce_ms, "nccl_serial_gather_ms": nccl_gather_ms,
                        "what": "per batch, inside the last step's event pair: k_pack_records (hit records of the batch's last "
                                "frame) + k_vertex_max + ONE snapshot of the accumulator block.  Then, on a side stream, path "
                                "'peer memory': ONE kernel per rank (k_peer_combine: flag barrier over NVLink, every rank folds "
                                "the snapshots of all ranks into its totals straight out of the peers' memory -- SUM over the "
                                "histogram words, MAX over the float bits of fmax | vmax --, rank 0 pulls every rank's records in "
                                "rank order; the counts are read on the device); path 'nccl': all_reduce SUM + ONE all_reduce MAX "
                                "of the snapshot, count exchange with a host read, all-gather of the records.  Batch b's "
                                "side-stream work overlaps batch b+1's frames; the last batch's is the exposed tail (inside the "
                                "timed total).  rank_spread_of_frames_ms: slowest minus fastest rank's own frames (the tail "
                                "includes waiting for the slowest rank).  serial_*: pack + snapshot + combine (e0..e1) and what "
                                "follows (e1..e2) run serially after the timed region; nccl_serial_*: the same work through the "
                                "NCCL path on the same ranks, for comparison",
                        "hit_records_gathered": gathered[0]},
            "checks": {"rays_per_frame": n_rays, "hits_per_frame": n_hits, "hist_total_all_ranks": hist_total,
                       "hits_reported_all_ranks": int(hits_reported[0]),
                       "hist_total_equals_hits": bool(hist_total == int(hits_reported[0]) and hist_total > 0)},
        }
        if sub:
            line["configs"] = sub
        if world == 1 and not args.no_configs:
            cfgs = {}
            for mesh in ("ns_1m", "c4_5m"):
                cfgs[mesh] = sub_config_dense(mesh, peak, flush)
            cfgs["c2_500k_build"] = {k: v for k, v in sub_config_dense("c2_500k", peak, flush, steps=4).items() if k.startswith("build") or k in ("triangles", "wide_nodes")}
            cfgs["c3_refit"] = sub_config_c3(flush)
            cfgs["c1"] = sub_config_c1(flush)
            cfgs["north_star_target_met"] = bool(cfgs["ns_1m"]["trace_mrays_s"] >= 1000.0)
            line["configs"] = cfgs
        if world == 1 and not args.no_cpu:
            from oracle import oracle as orc
            orc.build()
            cb = cpu_frames(V, F, K, H, W, 12.0, 40)
            line["cpu_baseline"] = {
                "value": cb["rays"] / cb["seconds"] / 1e6, "unit": "Mrays/s", "cores": cb["cores"], "kind": cb["kind"],
                "sample": f"{cb['frames']} full frames of the same workload ({cb['rays']} rays); per frame: float64 "
                          f"vertex posing + BVH build (per call, as the reference) + rays + closest hit + accumulation; "
                          f"cast-only {cb['rays'] / cb['cast_seconds'] / 1e6:.2f} Mrays/s; ms/frame {1e3 * cb['seconds'] / cb['frames']:.1f}"}
        print(json.dumps(line), flush=True)
    proj.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_strong(args):
    """--scaling strong: the headline line itself is the single-frame ray-sharded projection (total work fixed)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    sampler = ClockSampler(local)
    sampler.start()
    r = strong_scaling(args.mesh, rank, world, local, args.steps, flush)
    clocks = sampler.stop()
    if rank == 0:
        cfg = config_dict(args, r["triangles"], 1024, 1024)
        cfg["workload"] += "; ONE frame per step, its compacted ray list split over the ranks (strong scaling)"
        print(json.dumps({"metric": "Mrays/s", "value": r["mrays_s"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                          "warmup": 3, "ms_per_step": r["ms_per_frame"], "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
                          "gpu_launches": 5 * args.steps, "strong": r}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", default="c2_500k")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batches", type=int, default=2, help="combines per run of K steps (the last one is exposed)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (sub-results)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.scaling == "strong":
        run_strong(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
