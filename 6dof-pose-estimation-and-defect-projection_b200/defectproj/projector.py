"""Batched / multi-GPU surface: one process per GPU, mesh + BVH replicated.

The data path has no collective: every rank projects its own frames (`Projector.project_batch`, frames sharded) or
its own block of one frame's compacted ray list (`Projector.project_frame_sharded`, rays sharded) against its own
copy of the BVH.  Only the results are combined, once per batch:
    hist  int32  [nF]              all-reduce SUM   } on a SNAPSHOT of the accumulator block, on a side stream, while
    fmax | vmax  float32 bits      ONE all-reduce MAX } the next batch's frames run; the live accumulators stay local
    hits  variable length          counts exchanged once, then the records travel UNPADDED (to every rank, or to one)
Integer sums and float maxima do not depend on the order of combination, so an N-GPU result is bit-identical to the
1-GPU result.
"""
from __future__ import annotations

import numpy as np

from .core import Context, DevView

__all__ = ["Projector", "FrameStream", "BatchCombiner", "PeerCombiner", "shard_range", "combine_accumulators", "combine_block", "fold_block",
           "gather_hits", "gather_slices"]


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous block partition: rank r owns [lo, hi).  Blocks differ by at most one item and
    keep the global (frame-major, then row-major) order when concatenated by rank."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist
    return None


def _global_rank(dist, group, r):
    return dist.get_global_rank(group, r) if group is not None else r


def combine_accumulators(hist, fmax, vmax, group=None):
    """In-place all-reduce of three separate accumulator tensors (any device).  Callers that hold the library's live
    accumulators should NOT use this on them (a following accumulation would count earlier batches `world` times):
    use BatchCombiner, which reduces a snapshot."""
    dist = _dist(group)
    if dist is None:
        return hist, fmax, vmax
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(fmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX, group=group)
    return hist, fmax, vmax


def combine_block(block, max_from, group=None):
    """In-place combination of one accumulator block (int32 tensor: hist | fmax bits | vmax bits, see
    dp_accum_layout): SUM over the words [0, max_from), ONE MAX over [max_from, end) -- the bit patterns of
    non-negative floats order like integers, so fmax and vmax share a single collective."""
    dist = _dist(group)
    if dist is None:
        return block
    dist.all_reduce(block[:max_from], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(block[max_from:], op=dist.ReduceOp.MAX, group=group)
    return block


def fold_block(total, snapshot, max_from, group=None):
    """One batch into the running totals: the snapshot (this rank's accumulator block since the previous snapshot)
    is combined over the ranks in place, then total[:max_from] += it and total[max_from:] = max(., it)."""
    import torch
    combine_block(snapshot, max_from, group)
    total[:max_from].add_(snapshot[:max_from])
    torch.maximum(total[max_from:], snapshot[max_from:], out=total[max_from:])
    return total


def gather_hits(records, group=None, dst=None, counts=None):
    """Variable-length per-rank records ([n_r, C] tensor) -> [sum n_r, C] in rank order.  The counts are exchanged once
    (one small all-gather and one host read), then ONE all-gather moves every rank's rows: straight into place when the
    counts are equal (dense frames), otherwise padded to the largest count only (not to the buffers' capacity) and
    compacted by one concatenation.  dst=None: every rank returns the whole; dst=r: only rank r does (the viewer's rank),
    the others return None.  `counts` (list of ints) skips the exchange when the caller already knows them.
    (Measured on 2 B200: 12.6 MB per rank in 0.05 ms by all-gather; the point-to-point variant of this function showed
    millisecond outliers from lazily set-up channels and was dropped.)"""
    import torch
    dist = _dist(group)
    if dist is None:
        return records
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    records = records.contiguous()
    if counts is None:
        n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
        counts_t = torch.empty(world, dtype=torch.int64, device=records.device)
        dist.all_gather_into_tensor(counts_t, n, group=group)
        counts = counts_t.tolist()                              # the one host synchronisation of the gather
    m = max(counts) if counts else 0
    tail = tuple(records.shape[1:])
    if m == 0:
        return records[:0] if dst is None or rank == dst else None
    if records.shape[0] == m:
        pad = records
    else:
        pad = torch.empty((m,) + tail, dtype=records.dtype, device=records.device)
        pad[:records.shape[0]] = records
    buf = torch.empty((world * m,) + tail, dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    if dst is not None and rank != dst:
        return None
    if all(c == m for c in counts):
        return buf
    return torch.cat([buf[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)


def gather_slices(full, ranges, group=None, mine=None, dst=None):
    """`full`: a tensor every rank holds at full length; rank r owns rows ranges[r] = (lo, hi) of it (already filled,
    or given as `mine`).  dst=None: after the call every rank holds every slice -- ONE all-gather, in place, when the
    slices are equal-sized and tile the tensor's head (the dense-frame case), otherwise one broadcast per non-empty
    slice.  dst=r: only rank r receives the others' slices (one grouped point-to-point exchange)."""
    dist = _dist(group)
    if dist is None:
        if mine is not None:
            full[ranges[0][0]:ranges[0][1]].copy_(mine)
        return full
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if mine is not None:
        full[ranges[rank][0]:ranges[rank][1]].copy_(mine)
    if dst is not None:
        ops = []
        if rank == dst:
            for r, (lo, hi) in enumerate(ranges):
                if r != rank and hi > lo:
                    ops.append(dist.P2POp(dist.irecv, full[lo:hi], _global_rank(dist, group, r), group))
        elif ranges[rank][1] > ranges[rank][0]:
            ops.append(dist.P2POp(dist.isend, full[ranges[rank][0]:ranges[rank][1]], _global_rank(dist, group, dst), group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return full
    size = ranges[0][1] - ranges[0][0]
    equal = size > 0 and all(lo == r * size and hi == (r + 1) * size for r, (lo, hi) in enumerate(ranges))
    if equal and full.is_contiguous():
        dist.all_gather_into_tensor(full[:world * size], full[rank * size:(rank + 1) * size], group=group)
        return full
    for r, (lo, hi) in enumerate(ranges):
        if hi > lo:
            dist.broadcast(full[lo:hi], _global_rank(dist, group, r), group=group)
    return full


_DevView = DevView


class BatchCombiner:
    """Per-batch combination of the accumulators without touching the live ones.

    submit(): on the compute stream the per-vertex maxima are brought up to date, the whole accumulator block
    (hist | fmax | vmax, ONE allocation) is copied into a staging buffer and -- by default -- zeroed for the next
    batch; a side stream then all-reduces the snapshot (one SUM, one MAX) and folds it into the running totals
    (hist += , fmax/vmax = max).  The next batch's kernels run meanwhile.  result() waits for the side stream and
    returns the totals over every submitted batch and every rank.  Because only snapshots (deltas) are reduced, a
    histogram accumulated over several batches counts every hit exactly once.
    """

    def __init__(self, ctx: Context, group=None):
        import torch
        self.ctx, self.group = ctx, group
        self.dev = torch.device(f"cuda:{ctx.device}")
        base, h_off, f_off, v_off, nbytes = ctx.accum_layout()
        self.words, self.max_from = nbytes // 4, f_off // 4
        self.f_off, self.v_off = f_off // 4, v_off // 4
        self.live = torch.as_tensor(_DevView(base, self.words, "<i4"), device=self.dev)
        self.stage = [torch.empty(self.words, dtype=torch.int32, device=self.dev) for _ in range(2)]
        self.total = torch.zeros(self.words, dtype=torch.int32, device=self.dev)
        self.side = torch.cuda.Stream(device=self.dev)
        self.ev_snap = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.busy = [False, False]
        self.k = 0

    def reset_totals(self, stream=None):
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.side)
        with torch.cuda.stream(stream):
            self.total.zero_()

    def submit(self, stream=None, reset=True):
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        k = self.k
        self.k ^= 1
        if self.busy[k]:
            stream.wait_event(self.ev_done[k])              # the staging buffer's previous reduction has finished
        self.ctx.accum_flush(stream)                        # vmax derived from fmax (one small kernel)
        with torch.cuda.stream(stream):
            self.stage[k].copy_(self.live)
            if reset:
                self.live.zero_()
            self.ev_snap[k].record(stream)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_snap[k])
            fold_block(self.total, self.stage[k], self.max_from, self.group)
            self.ev_done[k].record(self.side)
        self.busy[k] = True
        return self

    def result(self, stream=None):
        """(hist int32 [nF], fmax float32 [nF], vmax float32 [nV]) totals; the caller's stream waits for the side stream."""
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.side)
        nF, nV = self.ctx.nF, self.ctx.nV
        t = self.total
        return (t[:nF], t[self.f_off:self.f_off + nF].view(torch.float32), t[self.v_off:self.v_off + nV].view(torch.float32))


class PeerCombiner:
    """BatchCombiner over peer-mapped memory (csrc/peer.cu): per batch ONE kernel of the library -- a flag barrier over
    NVLink, then every rank folds the accumulator snapshots of all ranks into its totals straight out of the peers'
    memory, and the gathering rank pulls every rank's hit records into one array -- instead of two all-reduces, a count
    exchange with a host read-back and a padded all-gather.  Same interface and same (bit-identical) totals as
    BatchCombiner; the process group is only used once, to exchange the 64-byte window handles.

    Slots alternate per batch.  A slot (snapshot + records) may be rewritten once this rank's most recent combine has
    completed: it passed the barrier every peer enters only after finishing the combine before -- `acquire()` makes the
    compute stream wait for exactly that (it is long done when batches are longer than a combine)."""

    def __init__(self, ctx: Context, rank: int, world: int, group=None, record_rows: int = 0, row_words: int = 3,
                 result_rays: int = 0, local_contexts=None, handles=None):
        import torch
        self.ctx, self.group, self.rank, self.world = ctx, group, int(rank), int(world)
        self.dev = torch.device(f"cuda:{ctx.device}")
        self.row_words = int(row_words)
        base, h_off, f_off, v_off, nbytes = ctx.accum_layout()
        self.words, self.max_from = nbytes // 4, f_off // 4
        self.f_off, self.v_off = f_off // 4, v_off // 4
        self.handle = ctx.peer_export(int(record_rows) * self.row_words * 4, int(result_rays))
        if local_contexts is None and handles is None:
            import torch.distributed as dist
            mine = torch.frombuffer(bytearray(self.handle), dtype=torch.uint8).to(self.dev)
            allh = torch.empty(world * len(self.handle), dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            handles = bytes(allh.cpu().numpy())
        self._local = local_contexts
        if handles is not None:
            ctx.peer_open(self.rank, self.world, handles)
        self.total = torch.zeros(self.words, dtype=torch.int32, device=self.dev)
        self.side = torch.cuda.Stream(device=self.dev)
        self.ev_snap = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.last_done = None
        self.k = 0
        self._acquired = None
        self._rec = [None, None]

    def open_local(self):
        """second phase for contexts of one process: call after EVERY context has been constructed (exported)"""
        self.ctx.peer_open_local(self.rank, self._local)
        return self

    def reset_totals(self, stream=None):
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.side)
        with torch.cuda.stream(stream):
            self.total.zero_()

    def acquire(self, stream=None):
        """The slot of the batch that is ending; `stream` may write its records and snapshot from here on."""
        import torch
        if self._acquired is None:
            stream = stream or torch.cuda.current_stream(self.dev)
            if self.last_done is not None:
                stream.wait_event(self.last_done)
            self._acquired = self.k
            self.k ^= 1
        return self._acquired

    def records(self, slot):
        """(int32 [rows, row_words] record slot, int64 [1] row count) of the own window: pass them to
        Context.pack_records_device(out=..., count_async=...) so the batch's hit records are packed in place."""
        from . import _lib
        if self._rec[slot] is None:
            self._rec[slot] = (self.ctx.peer_tensor(_lib.DP_PEER_RECORDS, slot, self.row_words),
                               self.ctx.peer_tensor(_lib.DP_PEER_REC_COUNT, slot))
        return self._rec[slot]

    def submit(self, stream=None, reset=True, gather_root=-1, gathered=None, count_async=None):
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        k = self.acquire(stream)
        self._acquired = None
        self.ctx.peer_snapshot(k, reset, stream)               # vertex maxima + snapshot (+ zero) of the live block
        self.ev_snap[k].record(stream)
        self.side.wait_event(self.ev_snap[k])
        self.ctx.peer_combine(k, self.total, gather_root, gathered if self.rank == gather_root else None, count_async,
                              stream=self.side)
        self.ev_done[k].record(self.side)
        self.last_done = self.ev_done[k]
        return self

    def result(self, stream=None):
        import torch
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.side)
        nF, nV = self.ctx.nF, self.ctx.nV
        t = self.total
        return (t[:nF], t[self.f_off:self.f_off + nF].view(torch.float32), t[self.v_off:self.v_off + nV].view(torch.float32))

    def check(self):
        """raises if a wait of this rank timed out (a peer never arrived); synchronises"""
        e = self.ctx.peer_status()
        if e:
            raise RuntimeError(f"peer exchange: a wait on channel {e - 1} timed out (a rank did not reach the same call)")


class _LazyHits:
    """The shard's hit count of a ray-sharded frame: in pinned memory once the frame's kernels have run (int() waits)."""

    def __init__(self, cnt):
        self._cnt = cnt

    def __int__(self):
        import torch
        torch.cuda.current_stream().synchronize()
        return int(self._cnt[1])

    __index__ = __int__


class _LazyCount(_LazyHits):
    """counts[i] of a frame queued without waiting: int() synchronises the current stream"""

    def __init__(self, cnt, i):
        self._cnt, self._i = cnt, i

    def __int__(self):
        import torch
        torch.cuda.current_stream().synchronize()
        return int(self._cnt[self._i])

    __index__ = __int__


class Projector:
    """Mesh + BVH on this process's GPU; frames of a batch, or the rays of one frame, sharded over the process group."""

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
        self.combiner = BatchCombiner(self.ctx, group)
        self.peer = None
        self._cnt = None
        self._res_slot = 0
        self._records = self._gathered = self._rows = None
        import os
        if self.world > 1 and os.environ.get("DP_PEER", "1") != "0" and dist.get_backend(group) == "nccl":
            self.enable_peer()

    def enable_peer(self, record_rows: int = 0, row_words: int = 3, result_rays: int = 0):
        """(Re)create the exchange windows over peer-mapped memory (collective: every rank calls it with the same sizes).
        record_rows: capacity of a batch's hit records per rank; result_rays: rays of a ray-sharded frame.  When any rank
        cannot map its peers (no NVLink / PCIe peer access, ranks on several nodes) every rank keeps the NCCL combiner.
        Returns True when the peer path is in use."""
        import torch
        import torch.distributed as dist
        if self.world < 2:
            return False
        if self.peer is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)                   # nobody still reads a window that is about to go
            self.peer.ctx.peer_close()
            dist.barrier(group=self.group)                   # ... and nobody maps a window that is about to be freed
            self.peer = None
            self.combiner = BatchCombiner(self.ctx, self.group)
        ok, comb = 1, None
        try:
            comb = PeerCombiner(self.ctx, self.rank, self.world, self.group, record_rows, row_words, result_rays)
        except Exception as e:                               # DP_E_CUDA: no peer access / IPC unavailable
            ok, self.peer_error = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{self.device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag[0]) == 1:
            self.peer = self.combiner = comb
            return True
        if comb is not None:
            comb.ctx.peer_close()
        return False

    def close(self):
        """Collective when the peer path is in use: every rank unmaps its peers' windows before any window is freed."""
        import torch
        if self.peer is not None:
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self.ctx.peer_close()
            dist.barrier(group=self.group)
            self.peer = None
        self.ctx.close()

    def accumulators(self):
        """Torch views (no copy) of THIS rank's live hist int32 [nF], fmax float32 [nF], vmax float32 [nV]."""
        import torch
        self.ctx.accum_flush(torch.cuda.current_stream(self.device))
        h, f, v = self.ctx.accum_device_ptrs()
        dev = f"cuda:{self.device}"
        return (torch.as_tensor(_DevView(h, max(self.nF, 1), "<i4"), device=dev)[:self.nF],
                torch.as_tensor(_DevView(f, max(self.nF, 1), "<f4"), device=dev)[:self.nF],
                torch.as_tensor(_DevView(v, max(self.nV, 1), "<f4"), device=dev)[:self.nV])

    def combined(self):
        """(hist, fmax, vmax) over every batch submitted since the last reset and over all ranks."""
        return self.combiner.result()

    def project_batch(self, heat, K, poses, thr=0.5, mode="object", out=None, reduce=True, reset=True, records_to=None):
        """heat: CUDA tensor [B_local,H,W] holding THIS rank's frames (see shard_range);
        poses: [B_local,4,4] model->camera.  mode 'object' = one launch for the whole batch,
        'camera' = per-frame dp_pose_mesh (refit) + launch, the reference-literal arithmetic.
        reduce: the batch's accumulators are snapshotted, zeroed and combined over the ranks into the running
        totals (`combined()`), asynchronously; reset: the totals start from zero with this batch (False: they keep
        accumulating -- every hit is still counted once).  records_to=r (mode 'object', `out` with 'pixel', 't_hit' and
        'face'): the batch's hit records -- rows (pixel, t_hit bits, face) of the rays that hit, the rays the reference
        keeps (/root/reference/src/defect_projection.py:259-264) -- of ALL ranks are gathered on rank r in rank order
        (`hit_records()`): inside the combine kernel on the peer path (enable_peer(record_rows=...) sizes the windows),
        by one count exchange + one all-gather on the NCCL path.  Returns (n_rays, n_hits) of this rank."""
        import torch
        if records_to is not None and (mode != "object" or not reduce or out is None or any(k not in out for k in ("pixel", "t_hit", "face"))):
            raise ValueError("records_to needs mode='object', reduce=True and out with 'pixel', 't_hit' and 'face'")
        if reset:
            self.ctx.accum_reset(torch.cuda.current_stream())
            self.combiner.reset_totals()
        K = np.asarray(K, np.float64).reshape(-1, 9)
        poses = np.asarray(poses, np.float64).reshape(-1, 4, 4)
        self.ctx.set_ray_shard(0, 1)             # whole frames (a no-op unless project_frame_sharded ran before)
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
        if records_to is not None:
            self._records = self._gather_records(out, n, int(records_to))
        elif reduce:
            self.combiner.submit(reset=True)
        return n, h

    def _gather_records(self, out, n, root):
        import torch
        ctx = self.ctx
        if self.peer is not None:
            comb = self.peer
            k = comb.acquire()
            rec, cnt = comb.records(k)
            if rec is None or rec.shape[0] < n:
                raise MemoryError(f"the exchange windows hold {0 if rec is None else rec.shape[0]} records per rank, the batch has "
                                  f"{n} rays: enable_peer(record_rows=...) on every rank")
            ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n, out=rec, count_async=cnt, sync=False)
            rows_cap = self.world * rec.shape[0]
            if self.rank == root and (self._gathered is None or self._gathered.shape[0] < rows_cap):
                self._gathered = torch.empty((rows_cap, rec.shape[1]), dtype=torch.int32, device=rec.device)
            if self._rows is None:
                self._rows = torch.zeros(1, dtype=torch.int64).pin_memory()
            comb.submit(reset=True, gather_root=root, gathered=self._gathered, count_async=self._rows)
            return ("peer", root)
        rec = ctx.pack_records_device(out["t_hit"], out["face"], pixel=out["pixel"], n=n)
        self.combiner.submit(reset=True)
        return ("nccl", root, gather_hits(rec, group=self.group, dst=root))

    def hit_records(self):
        """The records gathered by the last project_batch(records_to=r): int32 CUDA tensor [m, 3] (pixel, t_hit bits, face)
        on rank r, None elsewhere.  Waits for the combine."""
        import torch
        if self._records is None:
            return None
        if self._records[0] == "nccl":
            return self._records[2]
        torch.cuda.current_stream().wait_stream(self.peer.side)
        torch.cuda.current_stream().synchronize()               # the row count is in pinned memory now
        self.peer.check()
        return self._gathered[:int(self._rows[0])] if self.rank == self._records[1] else None

    def project_frame_sharded(self, heat, K, pose, thr=0.5, out=None, gather="all", reduce=True, reset=True,
                              gather_stream=None):
        """ONE frame, its compacted ray list split over the ranks (the reference casts the frame's rays in
        one call, /root/reference/src/defect_projection.py:247-256).  Every rank holds the same `heat` [H,W] (CUDA) and
        pose, compacts the whole frame and traces block `rank` of `world`.  `out`: full-length per-ray CUDA tensors
        ('t_hit', 'face', 'point', ...).  gather: "all" -- every rank ends up with every rank's slice (row-major order,
        bit-identical to the 1-GPU arrays; ONE all-gather per array when the shards are equal); "root" -- only rank 0
        does (the viewer's rank); None -- the slices stay where they are.  reduce: the accumulators' snapshot is
        combined into the running totals (asynchronously; callers that project many frames pass reduce=False and call
        `combiner.submit()` once per batch).  Returns (n_rays of the frame, the hit count of THIS rank's shard -- an object
        whose int() waits for the frame --, (lo, hi) of this rank); the frame's hit total is the combined histogram's sum.
        Nothing here waits for the traversal: the call returns with the gathers queued behind it -- on `gather_stream`
        when given, so that a sequence of frames overlaps frame i's gathers with frame i+1's kernels."""
        import torch
        ctx = self.ctx
        if reset:
            ctx.accum_reset(torch.cuda.current_stream())
            self.combiner.reset_totals()
        if heat.dim() == 2:
            heat = heat[None]
        H, W = heat.shape[-2:]
        # the shard stays set between frames (changing it drops the learnt packet schedule); project_batch restores (0, 1)
        ctx.set_ray_shard(self.rank, self.world)
        if gather == "peer":
            # the traversal itself stores this rank's slice into the result arrays of EVERY rank (NVLink stores from the
            # kernel's epilogue) and the call ends with a flag barrier: no collective, no host wait for the ray count.
            # `out` (a dict) receives views of this rank's result slot: whole-frame t_hit / face (/ point) once the
            # stream has passed the call.  Slots alternate per frame (a frame's arrays live until the frame after next).
            from . import _lib
            if self.peer is None or self.world < 2:
                raise RuntimeError("gather='peer' needs enable_peer(result_rays=...) on every rank")
            slot = self._res_slot
            self._res_slot ^= 1
            with_points = out is not None and "point" in out
            ctx.peer_results(slot, with_points)
            if self._cnt is None:
                self._cnt = torch.zeros(2, dtype=torch.int64).pin_memory()
            ctx.project_device(heat, K, np.asarray(pose, np.float64).reshape(1, 4, 4), thr, "object", True,
                               out={"counts": self._cnt}, sync=False)
            if out is not None:
                out["t_hit"] = ctx.peer_tensor(_lib.DP_PEER_T_HIT, slot)
                out["face"] = ctx.peer_tensor(_lib.DP_PEER_FACE, slot)
                if with_points:
                    out["point"] = ctx.peer_tensor(_lib.DP_PEER_POINT, slot)
            if reduce:
                self.combiner.submit(reset=True)
            return _LazyCount(self._cnt, 0), _LazyCount(self._cnt, 1), None
        if self.peer is not None:
            ctx.peer_results(-1)                 # the caller's `out` tensors receive the results (collective gathers below)
        # the ray count arrives in pinned memory straight from the compaction kernel: the host sizes and queues the slice
        # gathers while the traversal still runs, and never waits for the frame
        if self._cnt is None:
            self._cnt = torch.zeros(2, dtype=torch.int64).pin_memory()
        cnt = self._cnt
        cnt[0] = -1
        o = dict(out) if out else {}
        o["counts"] = cnt
        ctx.project_device(heat, K, np.asarray(pose, np.float64).reshape(1, 4, 4), thr, "object", True, out=o, sync=False)
        cview = cnt.numpy()
        while cview[0] < 0:
            pass
        n = int(cview[0])
        ranges = [Context.shard_slots(r, self.world, n, H, W) for r in range(self.world)]
        if gather and out and _dist(self.group) is not None:
            cur = torch.cuda.current_stream()
            gs = gather_stream if gather_stream is not None else cur
            if gs is not cur:
                ev = torch.cuda.Event()
                ev.record(cur)
                gs.wait_event(ev)                                   # the gathers follow this frame's kernels ...
            with torch.cuda.stream(gs):                             # ... on a stream of their own when asked: the next
                for k, t in out.items():                            # frame's kernels overlap them (caller double-buffers `out`)
                    if k not in ("counts", "pixel", "intensity"):   # the selection is replicated: only results travel
                        gather_slices(t, ranges, group=self.group, dst=0 if gather == "root" else None)
        if reduce:
            self.combiner.submit(reset=True)
        return n, _LazyHits(cnt), ranges[self.rank]


class FrameStream:
    """Host-buffer streaming of frames through one context with copies and kernels overlapped.

    Three CUDA streams: H2D of frame i+1, kernels of frame i and D2H of frame i-1 run concurrently on `nbuf` sets of
    device buffers (two; a third so that the kernels never wait for a read-back measured equal: on the bench frame the
    kernel stream, not the read-back, is the longer chain -- scripts/stream_timeline2.py).  A frame's read-back is queued right behind its
    kernels, sized like the last frame whose counts are known (`speculate`; its real counts are checked before its
    device buffers are reused, and rows a larger frame still misses are fetched then); without a prediction the host
    waits for the frame's 16-byte counts and queues the exact-size copy.  A dense frame (every pixel selected) does not
    ship its pixel list: it is the identity, handed out as one shared read-only array.  Inputs and outputs are pinned host
    tensors, so every copy is a real DMA.  Results are identical to Context.project on the same frame.
    """

    _DT = {"pixel": "int32", "intensity": "float32", "t_hit": "float32", "face": "int32", "point": "float32"}

    def __init__(self, ctx: Context, H: int, W: int, want=("pixel", "t_hit", "face"), cap=None, ring: int = 3,
                 heat_dtype="float32", nbuf: int = 2):
        import torch
        self.ctx, self.H, self.W, self.want = ctx, int(H), int(W), tuple(want)
        self.cap = int(cap) if cap is not None else self.H * self.W
        self.ring = max(3, int(ring))
        dev = f"cuda:{ctx.device}"
        self.dev = dev
        hd = getattr(torch, heat_dtype)
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.nbuf = max(2, int(nbuf))                  # device-side slots (heatmap, results, counts) a frame cycles through
        self.heat_d = [torch.empty((1, self.H, self.W), dtype=hd, device=dev) for _ in range(self.nbuf)]

        def outs(device, pin):
            d = {}
            for k in self.want:
                shp = (self.cap, 3) if k == "point" else (self.cap,)
                t = torch.empty(shp, dtype=getattr(torch, self._DT[k]), device=device)
                d[k] = t.pin_memory() if pin else t
            return d
        self.out_d = [outs(dev, False) for _ in range(self.nbuf)]
        self.out_h = [outs("cpu", True) for _ in range(self.ring)]
        self.counts_h = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(self.nbuf)]
        self._identity = None
        self.last_d2h_bytes = 0
        self.speculate = True        # queue a frame's read-back behind its kernels, sized like the previous frame (see run)

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
        NB = self.nbuf
        ev_in = [torch.cuda.Event() for _ in range(NB)]
        ev_k = [torch.cuda.Event() for _ in range(NB)]
        ev_out = [torch.cuda.Event() for _ in range(self.ring)]
        done_k = [False] * NB
        pending = []                                   # frames whose D2H has been queued, not yet yielded
        next_d2h = 0
        self._pred = None                              # (rays, dense) of the last frame handed out
        prof = [] if getattr(self, "profile", False) else None     # per frame: timing events around each stage

        def mark(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e

        def copy_out(i, m, dense, lo=0):
            """D2H of rows [lo, m) of frame i's results on the output stream (the pixel list of a dense frame is not shipped)"""
            b, r = i % NB, i % self.ring
            for k in self.want:
                if k == "pixel" and dense:
                    continue                               # not copied: pop_result hands out the identity
                self.out_h[r][k][lo:m].copy_(self.out_d[b][k][lo:m], non_blocking=True)

        def queue_d2h(i, pred=None):
            """pred=None: wait for frame i's counts on the host, then queue its exact-size read-back.  pred=(rays, dense)
            of an earlier frame: queue the read-back NOW, behind frame i's kernels on the device (no host round trip between
            the kernels and the copy), sized by the prediction; validate() checks the real count before the frame's device
            buffers are reused and fetches what a larger frame still misses."""
            b = i % NB
            r = i % self.ring
            e = {"i": i, "r": r, "n": None, "nh": None}
            if pred is None:
                ev_k[b].synchronize()                      # the 16-byte counts of frame i are on the host now
                e["n"], e["nh"] = int(self.counts_h[b][0]), int(self.counts_h[b][1])
                e["m"] = min(e["n"], self.cap)
                e["dense"] = e["n"] == self.H * self.W and e["m"] == e["n"]    # every pixel selected: the pixel list is 0..n-1
                self._pred = (e["n"], e["dense"])
            else:
                e["m"], e["dense"] = min(pred[0], self.cap), pred[1]
            with torch.cuda.stream(self.s_out):
                if pred is not None:
                    self.s_out.wait_event(ev_k[b])
                if prof is not None:
                    prof[i]["out0"] = mark(self.s_out)
                copy_out(i, e["m"], e["dense"])
                ev_out[r].record(self.s_out)
                if prof is not None:
                    prof[i]["out1"] = mark(self.s_out)
            pending.append(e)

        def validate(i):
            """Frame i was read back on a prediction: its real counts (ev_k[i % NB] still stands for frame i) decide whether
            rows are missing; they are fetched while the frame's device buffers are still intact."""
            e = next((p for p in pending if p["i"] == i), None)
            if e is None or e["n"] is not None:
                return
            b = i % NB
            ev_k[b].synchronize()
            e["n"], e["nh"] = int(self.counts_h[b][0]), int(self.counts_h[b][1])
            m_true = min(e["n"], self.cap)
            dense_true = e["n"] == self.H * self.W and m_true == e["n"]
            if m_true > e["m"] or (e["dense"] and not dense_true):
                with torch.cuda.stream(self.s_out):
                    if e["dense"] and not dense_true:
                        copy_out(i, m_true, False)                 # the pixel list after all
                    else:
                        copy_out(i, m_true, dense_true, lo=e["m"])
                    ev_out[e["r"]].record(self.s_out)
            e["m"], e["dense"] = m_true, dense_true
            self._pred = (e["n"], dense_true)

        def pop_result():
            validate(pending[0]["i"])
            e = pending.pop(0)
            i, r, n, nh, m, dense = e["i"], e["r"], e["n"], e["nh"], e["m"], e["dense"]
            ev_out[r].synchronize()
            self.last_d2h_bytes = 16 + sum(self.out_h[r][k][:m].numel() * self.out_h[r][k].element_size()
                                           for k in self.want if not (k == "pixel" and dense))
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
            b = i % NB
            with torch.cuda.stream(self.s_in):
                if done_k[b]:
                    self.s_in.wait_event(ev_k[b])      # kernels of frame i-nbuf have consumed this heat buffer
                if prof is not None:
                    prof.append({"in0": mark(self.s_in)})
                self.heat_d[b].copy_(heats[i].reshape(1, self.H, self.W), non_blocking=True)
                ev_in[b].record(self.s_in)
                if prof is not None:
                    prof[i]["in1"] = mark(self.s_in)
            if i >= NB:
                # the device result buffer b is free once frame i-NB's D2H is done (all of it: a read-back queued on a
                # prediction is checked against the frame's real counts first)
                validate(i - NB)
                self.s_k.wait_event(ev_out[(i - NB) % self.ring])
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
            # read-backs in frame order: speculative (no host wait) once an earlier frame's size is known, else exact
            pred = self._pred if self.speculate else None
            while next_d2h <= (i if pred is not None else i - 1):
                queue_d2h(next_d2h, pred)
                next_d2h += 1
            while len(pending) > (2 if pred is not None else 1):
                yield pop_result()
        while next_d2h < B:
            queue_d2h(next_d2h, None)
            next_d2h += 1
        t_end.record(self.s_out)
        while pending:
            yield pop_result()
        t_end.synchronize()
        self.last_elapsed_ms = t_start.elapsed_time(t_end)
        if prof is not None:
            # ms since the start of the run for (H2D begin, H2D end, kernels begin, kernels end, D2H begin, D2H end)
            self.timeline = [[t_start.elapsed_time(f[k]) for k in ("in0", "in1", "k0", "k1", "out0", "out1")] for f in prof]
