"""Latency of the collectives the combine uses, on this box (torchrun, NCCL): CUDA events, after warm-up."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj.projector import gather_hits, gather_slices
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1 << 20
rec = torch.randint(0, 1 << 30, (n, 3), dtype=torch.int32, device="cuda")
def t(fn, name, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0) / reps
    if rank == 0: print(f"{name}: device {e0.elapsed_time(e1) / reps:.3f} ms  wall {wall:.3f} ms", flush=True)
t(lambda: gather_hits(rec, dst=0), "gather_hits dst=0 (counts + p2p 12.6MB/rank)")
t(lambda: gather_hits(rec, dst=0, counts=[n] * world), "gather_hits dst=0, counts known (p2p only)")
t(lambda: gather_hits(rec), "gather_hits all ranks (counts + broadcasts)")
outb = torch.empty((world * n, 3), dtype=torch.int32, device="cuda")
t(lambda: dist.all_gather_into_tensor(outb, rec), "all_gather_into_tensor 12.6MB/rank")
cnt = torch.tensor([n], dtype=torch.int64, device="cuda"); cout = torch.empty(world, dtype=torch.int64, device="cuda")
t(lambda: (dist.all_gather_into_tensor(cout, cnt), cout.tolist()), "counts all_gather + tolist")
h = torch.zeros(500000, dtype=torch.int32, device="cuda")
t(lambda: dist.all_reduce(h), "all_reduce SUM 2MB")
m = torch.zeros(750000, dtype=torch.int32, device="cuda")
t(lambda: dist.all_reduce(m, op=dist.ReduceOp.MAX), "all_reduce MAX 3MB")
full = torch.zeros(n, dtype=torch.int32, device="cuda")
rng = [(r * n // world, (r + 1) * n // world) for r in range(world)]
t(lambda: gather_slices(full, rng), "gather_slices equal (in-place all_gather 4MB total)")
t(lambda: gather_slices(full, rng, dst=0), "gather_slices dst=0 (grouped p2p)")
one = torch.zeros(1, dtype=torch.int64, device="cuda")
t(lambda: dist.all_reduce(one), "all_reduce 8 B")
dist.destroy_process_group()
