"""SPair-shaped pairs (BASELINE.json configs[0]): one batched launch vs the per-pair path (kernels 1-3).

    python tools/spair_probe.py [--pairs 1024]
"""
import argparse
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1024)
a = ap.parse_args()
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
sp = mv.spair
base = [syn.spair_pair(i) for i in range(16)]
B = a.pairs
feats = torch.stack([base[i % 16]["feats"] for i in range(B)]).cuda()  # (B, 2, 768, 14, 14): 1.2 MB per pair
ki = torch.stack([base[i % 16]["kps_i"] for i in range(B)]).cuda()
kj = torch.stack([base[i % 16]["kps_j"] for i in range(B)]).cuda()
ts = torch.tensor([base[i % 16]["thresh_scale"] for i in range(B)]).cuda()
hits = torch.zeros(2, dtype=torch.int64, device="cuda")
for _ in range(2):
    sp.compute_errors_batch(feats, ki, kj, ts, 224, hits=hits)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sp.compute_errors_batch(feats, ki, kj, ts, 224, hits=hits)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byts = feats.numel() * 4
print(f"batched: {B} pairs in {ms:.3f} ms = {B / ms * 1e3:.0f} pairs/s; {byts / ms / 1e6:.0f} GB/s of feature reads "
      f"({100 * byts / ms / 1e6 / 6544:.0f} % of 6544 GB/s)")
n = 64
t0 = time.perf_counter()
for i in range(n):
    p = base[i % 16]
    sp.compute_errors_from_features(feats[i], p["kps_i"], p["kps_j"], p["thresh_scale"], 224, hits=hits)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"per-pair path (kernels 1-3, ~10 launches + 1 sync per pair): {n / dt:.0f} pairs/s")
