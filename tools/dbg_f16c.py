import importlib, sys, os, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
mv = importlib.import_module("midvision-probe_b200")
import test_gpu_f16c as T
L, C_ = mv._lib, mv.correspondence
for C in (768, 200):
    X = T.collinear_rows(300, C, 1) * 3.0
    mu = C_._center(X.cuda().contiguous(), 300)
    hi, lo, _ = T.f16c_rows(mv, X, L.MV_ROLE_TARGET, center=mu)
    ss = (X.cuda() * X.cuda()).sum(1, keepdim=True)
    y = X.cuda() * (1.0 / ss.sqrt().clamp(min=1e-12)) - mu[None]
    got = hi[:, :C].float()
    ulp = (y.abs().clamp(min=2.0 ** -14) * 2.0 ** -10)
    viol = (got - y).abs() - (0.5 * ulp * 1.01 + 1e-9)
    idx = viol.argmax()
    r, c = int(idx // C), int(idx % C)
    print(C, "max violation", viol.max().item(), "at", r, c, "y", y[r, c].item(), "got", got[r, c].item(), "y.half", y[r, c].half().float().item(),
          "n viol", int((viol > 0).sum()), "frac", float((viol > 0).float().mean()))
    # same without centre
    hi2, _, _ = T.f16c_rows(mv, X, L.MV_ROLE_TARGET, center=None)
    x = X.cuda() * (1.0 / ss.sqrt().clamp(min=1e-12))
    print("   no centre: mismatches vs torch half", int((hi2[:, :C].float() != x.half().float()).sum()), "max abs diff", (hi2[:, :C].float() - x).abs().max().item())
    print("   centred : mismatches vs torch half", int((got != y.half().float()).sum()), "of", got.numel())
