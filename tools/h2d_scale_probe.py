"""Aggregate host -> device bandwidth of one box when N ranks copy from pinned memory at the same time (the ceiling of
the end-to-end arm of bench.py at N GPUs).   torchrun --nproc-per-node N tools/h2d_scale_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 19_660_800 // 4 * 4  # one NAVI-shaped pair: 19.6 MB
src = [torch.empty(n // 4, dtype=torch.float32).pin_memory() for _ in range(8)]
for s in src:
    s.normal_()
dst = torch.empty(n // 4, dtype=torch.float32, device="cuda")
for s in src:
    dst.copy_(s, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 400
t0 = time.perf_counter()
for i in range(reps):
    dst.copy_(src[i % 8], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([reps * n / dt / 1e9], device="cuda")
if world > 1:
    dist.all_reduce(gbs)
if rank == 0:
    print(f"N={world}: aggregate pinned host -> device {gbs.item():.1f} GB/s ({gbs.item() / world:.1f} per GPU), cpus {os.cpu_count()}")
if world > 1:
    dist.destroy_process_group()
