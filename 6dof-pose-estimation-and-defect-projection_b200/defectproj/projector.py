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

__all__ = ["Projector", "shard_range", "combine_accumulators", "gather_hits"]


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
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    pad = torch.zeros((m,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    pad[:records.shape[0]] = records
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


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
