"""Batched / multi-GPU surface: one process per GPU, frames sharded, mesh + BVH replicated.

The data path has no collective: every rank projects its own frames against its own copy of
the BVH.  Only the results are combined:
    hist  int32  [nF]  all-reduce SUM        fmax / vmax  float32  all-reduce MAX
    hits  variable length -> all-gather of counts, then padded all-gather of the records
Integer sums and float maxima do not depend on the order of combination, so an N-GPU result
is bit-identical to the 1-GPU result.
"""
from __future__ import annotations

import numpy as np

from .core import Context

__all__ = ["Projector", "FrameStream", "shard_range", "combine_accumulators", "gather_hits"]


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous block partition: rank r owns [lo, hi).  Blocks differ by at most one item and
    keep the global (frame-major, then row-major) order when concatenated by rank."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_accumulators(hist, fmax, vmax, group=None):
    """In-place all-reduce of the three accumulators (torch tensors on any device)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hist, fmax, vmax
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(fmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX, group=group)
    return hist, fmax, vmax


def gather_hits(records, group=None):
    """All-gather variable-length per-rank records ([n_r, C] tensor) -> [sum n_r, C] in rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    counts_t = torch.empty(world, dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts_t, n, group=group)
    counts = counts_t.tolist()                                  # the one host synchronisation of the gather
    m = max(counts) if counts else 0
    tail = tuple(records.shape[1:])
    if records.shape[0] == m:
        pad = records.contiguous()
    else:
        pad = torch.zeros((m,) + tail, dtype=records.dtype, device=records.device)
        pad[:records.shape[0]] = records
    buf = torch.empty((world * m,) + tail, dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(buf, pad, group=group)          # one flat collective, no per-rank staging copies
    if all(c == m for c in counts):
        return buf
    return torch.cat([buf[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)


class _DevView:
    """__cuda_array_interface__ view of library-owned device memory (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Projector:
    """Mesh + BVH on this process's GPU, frames of a batch sharded over the process group."""

    def __init__(self, V, F, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.ctx = Context(self.device)
        self.ctx.set_mesh(V, F)
        self.ctx.build_bvh()          # deterministic: every rank builds the identical BVH
        self.nV, self.nF = self.ctx.nV, self.ctx.nF

    def accumulators(self):
        """Torch views (no copy) of hist int32 [nF], fmax float32 [nF], vmax float32 [nV]."""
        import torch
        self.ctx.accum_flush(torch.cuda.current_stream(self.device))
        h, f, v = self.ctx.accum_device_ptrs()
        dev = f"cuda:{self.device}"
        return (torch.as_tensor(_DevView(h, max(self.nF, 1), "<i4"), device=dev)[:self.nF],
                torch.as_tensor(_DevView(f, max(self.nF, 1), "<f4"), device=dev)[:self.nF],
                torch.as_tensor(_DevView(v, max(self.nV, 1), "<f4"), device=dev)[:self.nV])

    def project_batch(self, heat, K, poses, thr=0.5, mode="object", out=None, reduce=True, reset=True):
        """heat: CUDA tensor [B_local,H,W] holding THIS rank's frames (see shard_range);
        poses: [B_local,4,4] model->camera.  mode 'object' = one launch for the whole batch,
        'camera' = per-frame dp_pose_mesh (refit) + launch, the reference-literal arithmetic.
        Returns (n_rays, n_hits) of this rank after synchronising."""
        import torch
        if reset:
            self.ctx.accum_reset(torch.cuda.current_stream())
        K = np.asarray(K, np.float64).reshape(-1, 9)
        poses = np.asarray(poses, np.float64).reshape(-1, 4, 4)
        if mode == "object":
            n, h = self.ctx.project_device(heat, K, poses, thr, "object", True, out=out, sync=True)
        else:
            n = h = 0
            for b in range(heat.shape[0]):
                self.ctx.pose_mesh(poses[b], torch.cuda.current_stream())
                kb = K[b if len(K) > 1 else 0]
                a, c = self.ctx.project_device(heat[b:b + 1], kb, None, thr, "camera", True, out=None, sync=True)
                n += a
                h += c
        if reduce and self.world > 1:
            combine_accumulators(*self.accumulators(), group=self.group)
        return n, h


class FrameStream:
    """Host-buffer streaming of frames through one context with copies and kernels overlapped.

    Three CUDA streams: H2D of frame i+1, kernels of frame i and D2H of frame i-1 run concurrently on
    double-buffered device memory; the host only waits for the 16-byte ray/hit counts of a frame before it
    queues that frame's (exact-size) result copy.  A dense frame (every pixel selected) does not ship its pixel
    list: it is the identity, handed out as one shared read-only array.  Inputs and outputs are pinned host tensors, so
    every copy is a real DMA.  Results are identical to Context.project on the same frame.
    """

    _DT = {"pixel": "int32", "intensity": "float32", "t_hit": "float32", "face": "int32", "point": "float32"}

    def __init__(self, ctx: Context, H: int, W: int, want=("pixel", "t_hit", "face"), cap=None, ring: int = 3,
                 heat_dtype="float32"):
        import torch
        self.ctx, self.H, self.W, self.want = ctx, int(H), int(W), tuple(want)
        self.cap = int(cap) if cap is not None else self.H * self.W
        self.ring = max(3, int(ring))
        dev = f"cuda:{ctx.device}"
        self.dev = dev
        hd = getattr(torch, heat_dtype)
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.heat_d = [torch.empty((1, self.H, self.W), dtype=hd, device=dev) for _ in range(2)]

        def outs(device, pin):
            d = {}
            for k in self.want:
                shp = (self.cap, 3) if k == "point" else (self.cap,)
                t = torch.empty(shp, dtype=getattr(torch, self._DT[k]), device=device)
                d[k] = t.pin_memory() if pin else t
            return d
        self.out_d = [outs(dev, False) for _ in range(2)]
        self.out_h = [outs("cpu", True) for _ in range(self.ring)]
        self.counts_h = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._identity = None
        self.last_d2h_bytes = 0

    def run(self, heats, K, poses, thr=0.5, frame="object", accumulate=True, before_kernels=None):
        """heats: sequence of pinned host tensors [H,W]; K [3,3] or per frame; poses [B,4,4].
        Yields (index, dict) per frame in order; the dict's arrays are views of a pinned ring buffer that is
        reused `ring` frames later.  `before_kernels(stream)` is called on the kernel stream ahead of every
        frame (bench.py uses it to flush L2).  After the generator is exhausted `last_elapsed_ms` holds the
        device time from the first H2D to the last D2H (CUDA events)."""
        import torch
        ctx = self.ctx
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(self.s_in)
        B = len(heats)
        K = np.asarray(K, np.float64).reshape(-1, 9)
        poses = None if poses is None else np.asarray(poses, np.float64).reshape(-1, 4, 4)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_k = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(self.ring)]
        done_k = [False, False]
        pending = []                                   # frames whose D2H has been queued, not yet yielded
        prof = [] if getattr(self, "profile", False) else None     # per frame: timing events around each stage

        def mark(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e

        def queue_d2h(i):
            b = i % 2
            ev_k[b].synchronize()                      # the 16-byte counts of frame i are on the host now
            n, nh = int(self.counts_h[b][0]), int(self.counts_h[b][1])
            m = min(n, self.cap)
            r = i % self.ring
            dense = n == self.H * self.W and m == n        # every pixel selected: the pixel list is 0..n-1
            with torch.cuda.stream(self.s_out):
                if prof is not None:
                    prof[i]["out0"] = mark(self.s_out)
                for k in self.want:
                    if k == "pixel" and dense:
                        continue                           # not copied: pop_result hands out the identity
                    self.out_h[r][k][:m].copy_(self.out_d[b][k][:m], non_blocking=True)
                ev_out[r].record(self.s_out)
                if prof is not None:
                    prof[i]["out1"] = mark(self.s_out)
            self.last_d2h_bytes = 16 + sum(self.out_h[r][k][:m].numel() * self.out_h[r][k].element_size()
                                           for k in self.want if not (k == "pixel" and dense))
            pending.append((i, r, n, nh, m))

        def pop_result():
            i, r, n, nh, m = pending.pop(0)
            ev_out[r].synchronize()
            dense = n == self.H * self.W and m == n
            res = {k: self.out_h[r][k][:m].numpy() for k in self.want if not (k == "pixel" and dense)}
            if "pixel" in self.want:
                if dense:
                    if self._identity is None:
                        self._identity = np.arange(self.H * self.W, dtype=np.uint32)
                        self._identity.flags.writeable = False
                    res["pixel"] = self._identity          # read-only, shared between dense frames
                else:
                    res["pixel"] = res["pixel"].view(np.uint32)
            res["n"], res["hits"] = n, nh
            if n > self.cap:
                raise MemoryError(f"frame {i}: {n} selected pixels exceed the stream capacity {self.cap}")
            return i, res

        for i in range(B):
            b = i % 2
            with torch.cuda.stream(self.s_in):
                if done_k[b]:
                    self.s_in.wait_event(ev_k[b])      # kernels of frame i-2 have consumed this heat buffer
                if prof is not None:
                    prof.append({"in0": mark(self.s_in)})
                self.heat_d[b].copy_(heats[i].reshape(1, self.H, self.W), non_blocking=True)
                ev_in[b].record(self.s_in)
                if prof is not None:
                    prof[i]["in1"] = mark(self.s_in)
            if i >= 2:
                # the device result buffer b is free once frame i-2's D2H is done
                self.s_k.wait_event(ev_out[(i - 2) % self.ring])
            self.s_k.wait_event(ev_in[b])
            if before_kernels is not None:
                with torch.cuda.stream(self.s_k):
                    before_kernels(self.s_k)
            if prof is not None:
                prof[i]["k0"] = mark(self.s_k)
            out = dict(self.out_d[b])
            out["counts"] = self.counts_h[b]
            ctx.project_device(self.heat_d[b], K[i if len(K) > 1 else 0], None if poses is None else poses[i:i + 1], thr,
                               frame, accumulate, out=out, sync=False, stream=self.s_k)
            ev_k[b].record(self.s_k)
            if prof is not None:
                prof[i]["k1"] = mark(self.s_k)
            done_k[b] = True
            if i >= 1:
                queue_d2h(i - 1)
            while len(pending) > 1:
                yield pop_result()
        if B:
            queue_d2h(B - 1)
        t_end.record(self.s_out)
        while pending:
            yield pop_result()
        t_end.synchronize()
        self.last_elapsed_ms = t_start.elapsed_time(t_end)
        if prof is not None:
            # ms since the start of the run for (H2D begin, H2D end, kernels begin, kernels end, D2H begin, D2H end)
            self.timeline = [[t_start.elapsed_time(f[k]) for k in ("in0", "in1", "k0", "k1", "out0", "out1")] for f in prof]
