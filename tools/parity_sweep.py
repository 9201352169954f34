"""Recall parity of the B200 path against the CPU oracle over several seeded pairs at BASELINE.json sizes.

    python tools/parity_sweep.py [--pairs 5]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restated  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=5)
a = ap.parse_args()
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
C_ = mv.correspondence
THR3, THR2 = [0.01, 0.02, 0.05], [5, 25, 50]
torch.set_num_threads(os.cpu_count() or 1)
worst = 0.0
for kind in ("navi", "scannet"):
    for i in range(2, 2 + a.pairs):
        if kind == "navi":
            p = syn.navi_pair(i)
            got = C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
            ref = restated.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
            Kmat = p["intrinsics"]
        else:
            p = syn.scannet_pair(i)
            got = C_.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 1000)
            ref = restated.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 1000)
            Kmat = p["K"]
        e3g, e2g = restated.pair_errors(got[0], got[1], p["Rt"], Kmat)
        e3r, e2r = restated.pair_errors(ref[0], ref[1], p["Rt"], Kmat)
        diffs = [abs(100.0 * (e3g < t).float().mean().item() - 100.0 * (e3r < t).float().mean().item()) for t in THR3]
        diffs += [abs(100.0 * (e2g < t).float().mean().item() - 100.0 * (e2r < t).float().mean().item()) for t in THR2]
        wdiff = float((got[2] - ref[2]).abs().max())
        worst = max(worst, max(diffs))
        print(f"{kind} pair {i}: max |recall diff| over 6 thresholds = {max(diffs):.3f} pp, max |weight diff| of the sorted matches = {wdiff:.2e}")
print(f"worst recall difference: {worst:.3f} pp (gate 0.1 pp)")
assert worst <= 0.1 + 1e-3
