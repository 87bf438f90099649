"""NCCL all-gather bandwidth probe (torchrun): the exchange step of the sharded ADI path in isolation."""
import os, time, datetime
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=120), device_id=torch.device(f"cuda:{local}"))
n, w = 79841, -(-246 // world)
pad = torch.randn(n, w, dtype=torch.float64, device="cuda")
out = torch.empty(world * n, w, dtype=torch.float64, device="cuda")
for _ in range(5):
    dist.all_gather_into_tensor(out, pad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_gather_into_tensor(out, pad)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
if rank == 0:
    print(f"all_gather {world} ranks, {out.numel() * 8 / 1e6:.0f} MB gathered per rank: {ms:.3f} ms  "
          f"({out.numel() * 8 * (world - 1) / world / ms / 1e6:.0f} GB/s received per rank)", flush=True)
dist.destroy_process_group()
