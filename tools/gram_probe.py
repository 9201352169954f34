"""GPU probe: how exact is a Gram matrix of unit source-pixel rows computed by kernel 2 (mv_k2_affinity, fp16 operands,
fp32 accumulation in TMEM) as HH + 2^-11 (HL + HL^T) from fp16 hi / lo planes, against an fp64 product?  And how long
does the launch take for the stacked (2 h w) x (2 h w) x C problem?   python tools/gram_probe.py"""
import importlib
import sys
from ctypes import c_size_t

import torch

sys.path.insert(0, ".")
mv = importlib.import_module("midvision-probe_b200")
bb = importlib.import_module("midvision-probe_b200.backbones")
L = mv._lib
C_ = mv.correspondence


def affinity(A, B, cluster=-1):
    n, C = A.shape
    m = B.shape[0]
    ld_s = (m + 3) // 4 * 4
    S = torch.empty((n, ld_s), dtype=torch.float32, device="cuda")
    rv = torch.empty((n, 2), dtype=torch.float32, device="cuda")
    ri = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    cb = torch.empty((m,), dtype=torch.int64, device="cuda")
    wsb = L.load().mv_k2_workspace_bytes(n, m)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")

    def run():
        L.call("mv_k2_affinity", L.ptr(A), C, L.ptr(B), C, n, m, C, None, None, L.MV_DTYPE_F16, cluster, L.ptr(S), ld_s,
               L.ptr(rv), L.ptr(ri), L.ptr(cb), L.ptr(ws), c_size_t(wsb), C_._stream())

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    return S[:, :m].clone(), e0.elapsed_time(e1) / 20 * 1e3


def probe(name, f0, f1):
    C, h, w = f0.shape
    s = torch.cat((f0.reshape(C, -1).t(), f1.reshape(C, -1).t())).cuda().float()
    U = torch.nn.functional.normalize(s, dim=1)
    hi = U.half()
    lo = ((U - hi.float()) * 2048.0).half()
    ref = U.double() @ U.double().t()
    for cl in (-1, 1, 2):
        HH, t_hh = affinity(hi.contiguous(), hi.contiguous(), cl)
        X, t_x = affinity(hi.contiguous(), lo.contiguous(), cl)
        G = HH.double() + (X.double() + X.double().t()) / 2048.0
        G32 = (HH + (X + X.t()) * (1.0 / 2048.0))
        exact_ops = hi.double() @ hi.double().t()  # what an error-free accumulation of the fp16 products would give
        e_hh = (HH.double() - exact_ops)
        e_g = G - ref
        e_g32 = G32.double() - ref
        print(f"{name} rows {s.shape[0]} C {C} cluster {cl}: HH {t_hh:.1f} us, X {t_x:.1f} us | accumulation error of HH: "
              f"max {e_hh.abs().max():.2e} rms {e_hh.pow(2).mean().sqrt():.2e} mean {e_hh.mean():+.2e} | "
              f"HH vs fp64 of the fp32 rows: max {(HH.double()-ref).abs().max():.2e} | "
              f"HH+X (fp64 combine): max {e_g.abs().max():.2e} rms {e_g.pow(2).mean().sqrt():.2e} mean {e_g.mean():+.2e} | "
              f"fp32 combine: max {e_g32.abs().max():.2e}")
    # the fp32 torch Gram for scale (what an fp32 reference itself gives)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = (U @ U.t()).double() - ref
    torch.backends.cuda.matmul.allow_tf32 = prev
    print(f"   torch fp32 matmul of the same rows: max {t32.abs().max():.2e} rms {t32.pow(2).mean().sqrt():.2e} mean {t32.mean():+.2e}")


if __name__ == "__main__":
    vit = bb.DenseViT(bb.vit_b16(0, img_size=224), multilayer=True).cuda()
    p = bb.navi_backbone_pair(0, vit, device="cuda", noise=0.7)
    probe("navi/vit", p["feat_0"], p["feat_1"])
    rn = bb.resnet50_layer4(0).cuda()
    p = bb.scannet_backbone_pair(0, rn, device="cuda", noise=1.0)
    probe("scannet/resnet", p["feat_0"], p["feat_1"])
