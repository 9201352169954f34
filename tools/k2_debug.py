"""Bring-up diagnostics for kernel 2 on a B200: run a few shapes, compare with fp32 torch on the device and
print where (which rows / columns / tiles) the results differ.  Not a test; prints and exits 0."""
import importlib
import os
import sys
import time
from ctypes import c_size_t

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mv = importlib.import_module("midvision-probe_b200")
L = mv._lib
C_ = mv.correspondence


def run(n, m, C, dtype, cluster, reps=0):
    g = torch.Generator().manual_seed(n * 7 + m)
    A = torch.randn(n, C, generator=g).cuda()
    B = torch.randn(m, C, generator=g).cuda()
    if dtype == "bf16":
        Ad, Bd = A.to(torch.bfloat16).contiguous(), B.to(torch.bfloat16).contiguous()
        Af, Bf = Ad.float(), Bd.float()
    else:
        Af = ((A.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()
        Bf = ((B.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()
        Ad, Bd = Af, Bf
    row_val = torch.zeros(n, 2, device="cuda")
    row_idx = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
    col_best = torch.zeros(m, dtype=torch.int64, device="cuda")
    wsb = L.load().mv_k2_workspace_bytes(n, m)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")

    def launch():
        L.call("mv_k2_sim_top2", L.ptr(Ad), L.ptr(Bd), n, m, C, None, None, 0 if dtype == "bf16" else 1, cluster,
               L.ptr(row_val), L.ptr(row_idx), L.ptr(col_best), L.ptr(ws), c_size_t(wsb), C_._stream())

    launch()
    torch.cuda.synchronize()
    col_val = torch.empty(m, device="cuda")
    col_idx = torch.empty(m, dtype=torch.int32, device="cuda")
    L.call("mv_k2_unpack_col", L.ptr(col_best), m, L.ptr(col_val), L.ptr(col_idx), C_._stream())
    torch.cuda.synchronize()
    if n * m <= 4e8:
        torch.backends.cuda.matmul.allow_tf32 = False
        S = Af @ Bf.t()
        k = min(2, m)
        val, idx = torch.topk(S, k, dim=1)
        ok1 = (row_idx[:, 0].long() == idx[:, 0])
        verr = (row_val[:, :k] - val).abs().max().item()
        cval, cidx = S.max(dim=0)
        okc = (col_idx.long() == cidx)
        print(f"[{dtype} mc={cluster}] n={n} m={m} C={C}: row top-1 match {ok1.float().mean().item():.4f} "
              f"max|dval| {verr:.3e}  col match {okc.float().mean().item():.4f} max|dcol| {(col_val - cval).abs().max().item():.3e}", flush=True)
        if ok1.float().mean().item() < 0.99:
            bad = (~ok1).nonzero().squeeze(1)[:8].tolist()
            for r in bad:
                print("   row", r, "got", row_idx[r].tolist(), [f"{x:.4f}" for x in row_val[r].tolist()], "want", idx[r].tolist(),
                      [f"{x:.4f}" for x in val[r].tolist()])
            print("   bad rows mod 128 histogram (8 bins):", torch.histc(((~ok1).nonzero().squeeze(1) % 128).float(), bins=8, min=0, max=128).tolist())
    if reps:
        for _ in range(3):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"[{dtype} mc={cluster}] n={n} m={m} C={C}: {ms:.3f} ms/launch  {2.0 * n * m * C / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    quick = [(128, 256, 64), (300, 280, 64), (1000, 777, 768), (2048, 2048, 256)]
    for dtype in ("bf16", "tf32"):
        for shp in quick:
            run(*shp, dtype, 0)
    for mc in (2, 4):
        for shp in quick[1:]:
            run(*shp, "bf16", mc)
    for mc in (0, 2, 4):
        run(19200, 19200, 768, "bf16", mc, reps=10)
    run(19200, 19200, 768, "tf32", 0, reps=5)
    run(19200, 19200, 2048, "bf16", 0, reps=5)
