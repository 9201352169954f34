"""Same-process A/B of kernel 2's two schedules (tile-granular vs stream-K) at the NAVI shape, kernel timed alone
(mv_k2_profile_*): back-to-back launches (the power-capped regime of the pipeline) and launches separated by idle
gaps (the regime of a synchronous helper call).   python tools/k2_streamk_ab.py [n m C]"""
import ctypes
import importlib
import sys
import time
from ctypes import c_size_t

import torch

sys.path.insert(0, ".")
mv = importlib.import_module("midvision-probe_b200")
L = mv._lib
C_ = mv.correspondence


def main():
    n, m, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (5024, 5024, 3080)
    lib = L.load()
    g = torch.Generator().manual_seed(0)
    ld = (C + 63) // 64 * 64
    A = torch.zeros(n, ld, dtype=torch.float16, device="cuda")
    B = torch.zeros(m, ld, dtype=torch.float16, device="cuda")
    A[:, :C] = torch.nn.functional.normalize(torch.randn(n, C, generator=g), dim=1).half().cuda()
    B[:, :C] = torch.nn.functional.normalize(torch.randn(m, C, generator=g), dim=1).half().cuda()
    rv = torch.empty((n, 2), dtype=torch.float32, device="cuda")
    ri = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    cb = torch.empty((m,), dtype=torch.int64, device="cuda")
    wsb = lib.mv_k2_workspace_bytes(n, m)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")

    def launch():
        L.call("mv_k2_sim_top2_ld", L.ptr(A), ld, L.ptr(B), ld, n, m, C, None, None, L.MV_DTYPE_F16, -1, L.ptr(rv), L.ptr(ri),
               L.ptr(cb), L.ptr(ws), c_size_t(wsb), C_._stream())

    def timed(reps, gap_s):
        lib.mv_k2_profile_begin(reps)
        for _ in range(reps):
            launch()
            if gap_s:
                torch.cuda.synchronize()
                time.sleep(gap_s)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * reps)()
        k = lib.mv_k2_profile_read(buf, reps)
        lib.mv_k2_profile_begin(0)
        v = sorted(buf[i] for i in range(k))
        return 1e3 * v[len(v) // 2], 1e3 * sum(v) / len(v)

    ref = {}
    for rnd in range(3):
        for sk in (1, 0):
            lib.mv_k2_set_streamk(sk)
            launch()
            torch.cuda.synchronize()
            ref[sk] = (ri.clone(), cb.clone())
            b2b = timed(200, 0.0)
            iso = timed(40, 0.002)
            fl = 2.0 * n * m * C
            print(f"round {rnd} streamk={sk}: back-to-back median {b2b[0]:.1f} us (mean {b2b[1]:.1f}) = {fl / b2b[0] / 1e6:.0f} TFLOP/s | "
                  f"isolated median {iso[0]:.1f} us (mean {iso[1]:.1f}) = {fl / iso[0] / 1e6:.0f} TFLOP/s")
    same = torch.equal(ref[0][0][:, 0], ref[1][0][:, 0])
    print("row arg-max identical between the schedules:", same, "| column records identical:", torch.equal(ref[0][1], ref[1][1]))


if __name__ == "__main__":
    main()
