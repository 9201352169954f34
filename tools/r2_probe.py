"""One pair of a dense workload through the eager (un-graphed) path a few times: the program the round-2 ncu captures
profile (`-k regex:<kernel>` picks the kernel).

    python tools/r2_probe.py [--workload navi|scannet] [--reps R] [--features backbone|gaussian] [--k1-grid]
"""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="navi")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--features", default="backbone")
ap.add_argument("--k1-grid", action="store_true")
ap.add_argument("--dtype", default="f16")
a = ap.parse_args()
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
bb = importlib.import_module("midvision-probe_b200.backbones")
C_, ev = mv.correspondence, mv.evaluation
C_.set_match_precision(dtype=a.dtype, k1_grid=int(a.k1_grid))
dev = torch.device("cuda")
if a.features == "gaussian":
    p = (syn.navi_pair if a.workload == "navi" else syn.scannet_pair)(0)
elif a.workload == "navi":
    p = bb.navi_backbone_pair(0, bb.DenseViT(bb.vit_b16(0), multilayer=True).to(dev), device=dev, noise=0.7)
else:
    p = bb.scannet_backbone_pair(0, bb.resnet50_layer4(0).to(dev), device=dev, noise=1.0)
p = {k: (v.to(dev) if torch.is_tensor(v) and v.dim() >= 3 else v) for k, v in p.items()}
thr = ([0.01, 0.02, 0.05], [5, 25, 50])
acc = ev.RecallAccumulator(*thr, device=dev)
for _ in range(a.reps):
    if a.workload == "navi":
        ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], p["intrinsics"], p["Rt"], 1000, acc, sync=False)
    else:
        ev.match_and_score_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"], p["Rt"], 1000, acc, sync=False)
torch.cuda.synchronize()
print(a.workload, acc.summary()["recall_3d"])
