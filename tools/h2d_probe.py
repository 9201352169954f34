"""Pinned host -> device copy bandwidth for the feature-map sizes of the e2e arm (PCIe floor of the helper call)."""
import torch
for mb in (9.6, 19.6, 64):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb:5.1f} MB pinned: {ms * 1e3:7.1f} us  {mb / ms:6.1f} GB/s")
