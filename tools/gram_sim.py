"""CPU simulation of the low-rank ("Gram") proposal form of kernel 2 (kernel 2g).

Upsampling is linear, so for interpolated rows x_i = sum_s W0[i,s] src0[s], y_j = sum_t W1[j,t] src1[t]:

    cos(x_i, y_j) = sum_t A[i,t] * B[j,t],   A[i,t] = x_i . s1_t / (|x_i| |s1_t|),   B[j,t] = W1[j,t] |s1_t| / |y_j|

with K = h*w source pixels instead of C channels.  This script rounds the operands as the tensor-core path would
(fp16 planes of the source maps for the Gram matrix, fp16 A' = A - c_i and fp16 B) and counts how many of the
reference's matches / how much recall survive when that product PROPOSES the two candidates and the exact fp32
distances of the reference decide.  Test infrastructure: uses oracle/restated.py.

    python tools/gram_sim.py navi 0 1      python tools/gram_sim.py scannet 0
"""
import importlib
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import restated  # noqa: E402

bb = importlib.import_module("midvision-probe_b200.backbones")


import os
EXACT_G = int(os.environ.get("EXACT_G", "0"))


def h16(x):
    return x.half().float()


def sim_pair(kind, seed, models, variant="c1"):
    t0 = time.time()
    if kind == "navi":
        p = bb.navi_backbone_pair(seed, models["vit"], device="cpu", noise=0.7)
        f0m, f1m = p["feat_0"], p["feat_1"]
        C, h, w = f0m.shape
        x0, f0, _, _ = restated.xyz_side(f0m, p["xyz_grid_0"])
        x1, f1, _, _ = restated.xyz_side(f1m, p["xyz_grid_1"])
        eye = torch.eye(h * w).reshape(h * w, h, w)
        W0 = restated.xyz_side(eye, p["xyz_grid_0"])[1]
        W1 = restated.xyz_side(eye, p["xyz_grid_1"])[1]
        Kmat = p["intrinsics"]
        thr3, thr2 = [0.01, 0.02, 0.05], [5, 25, 50]
    else:
        p = bb.scannet_backbone_pair(seed, models["resnet"], device="cpu", noise=1.0)
        f0m, f1m = p["feat_0"], p["feat_1"]
        C, h, w = f0m.shape
        x0, f0, _ = restated.depth_side(f0m, p["depth_0"], p["K"])
        x1, f1, _ = restated.depth_side(f1m, p["depth_1"], p["K"])
        eye = torch.eye(h * w).reshape(h * w, h, w)
        W0 = restated.depth_side(eye, p["depth_0"], p["K"])[1]
        W1 = restated.depth_side(eye, p["depth_1"], p["K"])[1]
        Kmat = p["K"]
        thr3, thr2 = [0.01, 0.02, 0.05, 0.10], [5, 10, 20, 30, 40, 50]
    n, m = f0.shape[0], f1.shape[0]
    # ---- reference
    i0, i1, wgt, d, idx, _ = restated.correspondences_ratio_test(f0, f1, 1000, return_all=True)
    e3r, e2r = restated.pair_errors(x0[i0], x1[i1], p["Rt"], Kmat)
    ref_pairs = set(zip(i0.tolist(), i1.tolist()))
    a = F.normalize(f0, dim=-1)
    b = F.normalize(f1, dim=-1)
    S_ref = a @ b.t()
    top = torch.topk(S_ref, 2, dim=1)
    gap = top.values[:, 0] - top.values[:, 1]
    t_ref = time.time() - t0
    # ---- low-rank form
    s0 = f0m.reshape(C, h * w).t().contiguous()
    s1 = f1m.reshape(C, h * w).t().contiguous()
    sc0 = 2.0 ** torch.floor(torch.log2(1.0 / s0.abs().max()))  # power-of-two scale into fp16's comfortable range
    sc1 = 2.0 ** torch.floor(torch.log2(1.0 / s1.abs().max()))
    H0, H1 = h16(s0 * sc0 * 256), h16(s1 * sc1 * 256)
    if EXACT_G:   # hi + lo planes: the Gram to fp32 accuracy
        G01, G00, G11 = s0 @ s1.t(), s0 @ s0.t(), s1 @ s1.t()
    else:
        G01 = (H0 @ H1.t()) / (sc0 * sc1 * 65536)    # fp16 operands, fp32 accumulate: the HH Gram
        G00 = (H0 @ H0.t()) / (sc0 * sc0 * 65536)
        G11 = (H1 @ H1.t()) / (sc1 * sc1 * 65536)
    nx = ((W0 @ G00) * W0).sum(1).clamp_min(1e-24).sqrt()   # |x_i| from the Gram
    ny = ((W1 @ G11) * W1).sum(1).clamp_min(1e-24).sqrt()
    ns1 = G11.diag().clamp_min(1e-24).sqrt()                 # |s1_t|
    A = (W0 @ G01) / (nx[:, None] * ns1[None, :])            # (n, hw)
    B = W1 * ns1[None, :] / ny[:, None]                      # (m, hw)
    chk = (A @ B.t() - S_ref).abs().max().item()
    Br = h16(B)
    beta = B.sum(1)   # the TRUE row sums: the rounding of B must not leak into the constant term
    out = {}
    for var in (["c1", "cmax"] if variant == "all" else [variant]):
        if var == "plain":
            S = h16(A) @ Br.t()
        elif var == "c1":   # per-row centre (one fp16 value), exact through the augmentation column c_i * beta_j
            c = h16(A.mean(1))
            S = h16(A - c[:, None]) @ Br.t() + c[:, None] * beta[None, :]
        elif var == "cmax":  # centre = the row's maximum: A' is smallest where the competitive columns have their taps
            c = h16(A.max(1).values)
            S = h16(A - c[:, None]) @ Br.t() + c[:, None] * beta[None, :]
        elif var == "c1s":   # c1 with A' as two fp16 planes (hi + lo): B rounding alone
            c = h16(A.mean(1))
            Ah = h16(A - c[:, None])
            Al = h16((A - c[:, None] - Ah) * 2048) / 2048
            S = (Ah + Al) @ Br.t() + c[:, None] * beta[None, :]
        else:               # per-row and per-column centre
            c = h16(A.mean(1))
            e = h16((A - c[:, None]).mean(0))
            gam = B @ e
            S = h16(A - c[:, None] - e[None, :]) @ Br.t() + c[:, None] * beta[None, :] + gam[None, :]
        err = (S - S_ref).abs()
        cand = torch.topk(S, 2, dim=1).indices               # proposals
        # the reference's exact fp32 distances of the two proposals decide (kernel 3)
        dd = 1 - F.cosine_similarity(b[cand], a[:, None, :], dim=-1)
        swap = dd[:, 1] < dd[:, 0]
        dd = torch.where(swap[:, None], dd.flip(1), dd)
        cand = torch.where(swap[:, None], cand.flip(1), cand)
        wg = restated.ratio_weights(dd)
        src, dst, _ = restated.topk_matches(wg, cand[:, 0], 1000)
        e3, e2 = restated.pair_errors(x0[src], x1[dst], p["Rt"], Kmat)
        common = len(ref_pairs & set(zip(src.tolist(), dst.tolist())))
        clear = gap > 1e-3
        bad = int((cand[clear, 0] != top.indices[clear, 0]).sum())
        set_same = float((torch.sort(cand, 1).values == torch.sort(top.indices, 1).values).all(1).float().mean())
        rec = [abs(float((e3 < t).float().mean() - (e3r < t).float().mean())) * 100 for t in thr3] + \
              [abs(float((e2 < t).float().mean() - (e2r < t).float().mean())) * 100 for t in thr2]
        out[var] = dict(err_max=err.max().item(), err_rms=err.pow(2).mean().sqrt().item(), common=common, nn_bad=bad,
                        top2_set_same=set_same, worst_pp=max(rec))
    print(f"{kind} seed {seed}: n={n} m={m} hw={h*w} C={C} gap>1e-3: {float((gap > 1e-3).float().mean()):.3f} "
          f"recall_ref@first={float((e3r < thr3[0]).float().mean())*100:.1f}% identity check {chk:.2e} "
          f"(ref {t_ref:.0f}s, total {time.time()-t0:.0f}s)")
    for k, v in out.items():
        print(f"   {k:6s} " + " ".join(f"{a}={b:.3g}" if isinstance(b, float) else f"{a}={b}" for a, b in v.items()))
    return out


if __name__ == "__main__":
    kind = sys.argv[1]
    seeds = [int(s) for s in sys.argv[2:]] or [0]
    torch.set_num_threads(8)
    models = {}
    if kind == "navi":
        models["vit"] = bb.DenseViT(bb.vit_b16(0, img_size=224), multilayer=True)
    else:
        models["resnet"] = bb.resnet50_layer4(0)
    for s in seeds:
        sim_pair(kind, s, models, "all")
