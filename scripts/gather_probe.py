import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200"))
from defectproj.projector import gather_hits
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1 << 20
rec = torch.randint(0, 1 << 30, (n, 3), dtype=torch.int32, device="cuda")
def t(fn, name, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {1e3 * (time.perf_counter() - t0) / reps:.3f} ms", flush=True)
t(lambda: gather_hits(rec), "gather_hits (list all_gather)")
outb = torch.empty((world * n, 3), dtype=torch.int32, device="cuda")
t(lambda: dist.all_gather_into_tensor(outb, rec), "all_gather_into_tensor 12.6MB/rank")
cnt = torch.tensor([n], dtype=torch.int64, device="cuda"); cout = torch.empty(world, dtype=torch.int64, device="cuda")
t(lambda: (dist.all_gather_into_tensor(cout, cnt), cout.tolist()), "counts all_gather + tolist")
t(lambda: rec[rec[:, 2] >= 0], "boolean select")
h = torch.zeros(500000, dtype=torch.int32, device="cuda")
t(lambda: dist.all_reduce(h), "all_reduce hist 2MB")
dist.destroy_process_group()
