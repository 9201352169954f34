"""Kernel 1 alone on one image (default NAVI-shaped: bicubic, C=3072, 28x28 -> 112x112, ~5k live points), for ncu.

    python tools/k1_probe.py [--kind navi|scannet] [--reps R] [--nosync]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="navi")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nosync", action="store_true")
a = ap.parse_args()
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
C_ = mv.correspondence
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if a.kind == "navi":
    p = syn.navi_pair(0)
    f, g = p["feat_0"].cuda(), p["xyz_grid_0"].cuda()
    run = lambda: C_.prepare_xyz_side(f, g, dev, sync=not a.nosync)
else:
    p = syn.scannet_pair(0)
    f, g = p["feat_0"].cuda(), p["depth_0"].cuda()
    Kh, Kinv = C_._host_mat(p["K"]), C_._host_mat(p["K"].inverse())
    run = lambda: C_.prepare_depth_side(f, g, Kh, Kinv, dev, sync=not a.nosync)
s = run()
n = int(s.n_dev.item())
C = f.shape[0]
times = []
for _ in range(a.reps):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = sorted(times)[len(times) // 2]
byts = C * f.shape[1] * f.shape[2] * 4 + n * C * (6 if s.rows16 is not None else 4)
print(f"{a.kind} side: n={n} C={C} whole prepare (compact + coords + transpose + kernel 1) {ms * 1e3:.1f} us; "
      f"kernel-1 algorithmic bytes {byts / 1e6:.1f} MB")
