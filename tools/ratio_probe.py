"""Kernel 3 (ratio / mutual) alone on NAVI- or ScanNet-sized rows: fp32 rows vs split (bf16 hi + lo) rows."""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=5024)
ap.add_argument("--C", type=int, default=3072)
a = ap.parse_args()
mv = importlib.import_module("midvision-probe_b200")
L = mv._lib
n, C = a.n, a.C
g = torch.Generator(device="cuda").manual_seed(0)
A32 = torch.nn.functional.normalize(torch.randn(n, C, device="cuda", generator=g), dim=1)
B32 = torch.nn.functional.normalize(torch.randn(n, C, device="cuda", generator=g), dim=1)
Ah, Bh = A32.to(torch.bfloat16), B32.to(torch.bfloat16)
Al, Bl = (A32 - Ah.float()).to(torch.bfloat16), (B32 - Bh.float()).to(torch.bfloat16)
# neighbours of a smooth map: query i -> rows near i (the regime of the real evaluations)
base = torch.arange(n, device="cuda")
idx = torch.stack(((base + 3) % n, (base + 4) % n), dim=1).to(torch.int32).contiguous()
d = torch.empty(n, 2, device="cuda")
w = torch.empty(n, device="cuda")
mu = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()


def run(split):
    s = ctypes.c_void_p(st.cuda_stream)
    if split:
        L.call("mv_k3_ratio_mutual_split", L.ptr(Ah), L.ptr(Al), L.ptr(Bh), L.ptr(Bl), C, None, n, L.ptr(idx), None, 1,
               L.ptr(d), L.ptr(w), L.ptr(mu), s)
    else:
        L.call("mv_k3_ratio_mutual", L.ptr(A32), L.ptr(B32), C, None, n, L.ptr(idx), None, 1, L.ptr(d), L.ptr(w), L.ptr(mu), s)


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for split in (False, True):
    with torch.cuda.stream(st):
        run(split)
        st.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            run(split)
            e1.record(st)
            st.synchronize()
            ts.append(e0.elapsed_time(e1))
    us = sorted(ts)[3] * 1e3
    print(f"ratio kernel {'split' if split else 'fp32 '} rows n={n} C={C}: {us:.1f} us (L2 flushed), {3 * n * C * 4 / us / 1e3:.0f} GB/s")
