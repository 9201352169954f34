"""Kernel 2 alone on one shape (default: BASELINE.json configs[4], 19200 x 19200 x 768 bf16), for ncu:

    python tools/k2_stress.py [--n N --m M --C C --dtype bf16|tf32 --cluster 0|2|4 --reps R --variant iid|upsampled]
"""
import argparse
import importlib
import os
import sys
from ctypes import c_size_t

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=19200)
ap.add_argument("--m", type=int, default=19200)
ap.add_argument("--C", type=int, default=768)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--cluster", type=int, default=-1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variant", default="iid")
a = ap.parse_args()

mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
L, C_ = mv._lib, mv.correspondence
A, B = syn.stress_rows(0, n=a.n, m=a.m, C=a.C, variant=a.variant)
tf32 = a.dtype == "tf32"
Ad = (A.cuda() if tf32 else A.cuda().to(torch.bfloat16)).contiguous()
Bd = (B.cuda() if tf32 else B.cuda().to(torch.bfloat16)).contiguous()
rv = torch.empty(a.n, 2, device="cuda")
ri = torch.empty(a.n, 2, dtype=torch.int32, device="cuda")
cb = torch.empty(a.m, dtype=torch.int64, device="cuda")
wsb = L.load().mv_k2_workspace_bytes(a.n, a.m)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")


def launch():
    L.call("mv_k2_sim_top2", L.ptr(Ad), L.ptr(Bd), a.n, a.m, a.C, None, None, int(tf32), a.cluster, L.ptr(rv), L.ptr(ri),
           L.ptr(cb), L.ptr(ws), c_size_t(wsb), C_._stream())


for _ in range(2):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    launch()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"k2 {a.dtype} cluster={a.cluster} {a.n}x{a.m}x{a.C} {a.variant}: {ms:.4f} ms  {2.0 * a.n * a.m * a.C / ms / 1e9:.1f} TFLOP/s")
