"""Full-size parity on BACKBONE-shaped features: random-init ViT-B/16 (NAVI- and SPair-shaped configs) and
ResNet-50 layer4 (ScanNet-shaped config), several seeds, every operand type.

Real backbone maps are spatially smooth, low-rank and -- for the CNN -- all-positive and nearly collinear (mean
cosine between two pixels ~0.93), which is the regime that decides whether a reduced-precision tensor-core product
proposes the right two neighbours.  Gates (north star):
  * NN index identical to the reference on every row whose fp32 top-2 similarity gap exceeds 1e-3; the fraction of
    rows in that set is asserted against what DESIGN.md publishes;
  * recall within 0.1 percentage points of the reference at every threshold on the aggregate over the seeds (the
    quantity the benchmarks report), and within 0.2 pp on every single pair (two matches in 1000: on these features the
    reference's own fp32 brute-force k-NN disagrees with an fp64 search on ~8 % of the rows -- exact ties of the
    fp32-quantised 1 - cos -- so single matches flip between any two exact implementations).
The centred operand types ("f16": f16c rows, the default; "tf32": tf32c rows) must pass everywhere; plain bf16 passes
on the ViT features and is REPORTED on the ResNet features, where it is outside the gate by ~9 pp -- the reason it is
not the default."""
import importlib

import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu

SEEDS = [0, 1, 2, 3]
NAVI_THR3, NAVI_THR2 = [0.01, 0.02, 0.05], [5, 25, 50]
# the JSON thresholds of render_scannet_correspondence.py:129-147 (px) and :253-264 (cm)
SCAN_THR3, SCAN_THR2 = [0.01, 0.02, 0.05, 0.10], [5, 10, 20, 30, 40, 50]


@pytest.fixture(scope="module")
def bb():
    return importlib.import_module("midvision-probe_b200.backbones")


@pytest.fixture(scope="module")
def models(bb):
    return {"vit_multi": bb.DenseViT(bb.vit_b16(0, img_size=224), multilayer=True).cuda(),
            "vit_last": bb.DenseViT(bb.vit_b16(0, img_size=224), multilayer=False).cuda(),
            "resnet": bb.resnet50_layer4(0).cuda()}


_CACHE = {}


def oracle_case(kind, seed, bb, models):
    """inputs + everything the reference computes for them, once per (kind, seed)."""
    key = (kind, seed)
    if key in _CACHE:
        return _CACHE[key]
    if kind == "navi":
        p = bb.navi_backbone_pair(seed, models["vit_multi"], device="cuda", noise=0.7)
        x0, f0, _, _ = restated.xyz_side(p["feat_0"], p["xyz_grid_0"])
        x1, f1, _, _ = restated.xyz_side(p["feat_1"], p["xyz_grid_1"])
        Kmat = p["intrinsics"]
    else:
        p = bb.scannet_backbone_pair(seed, models["resnet"], device="cuda", noise=1.0)
        x0, f0, _ = restated.depth_side(p["feat_0"], p["depth_0"], p["K"])
        x1, f1, _ = restated.depth_side(p["feat_1"], p["depth_1"], p["K"])
        Kmat = p["K"]
    i0, i1, w, d, idx, _ = restated.correspondences_ratio_test(f0, f1, 1000, return_all=True)
    e3, e2 = restated.pair_errors(x0[i0], x1[i1], p["Rt"], Kmat)
    o = restated.similarity_top2_and_mutual(f0, f1)
    out = {"p": p, "f0": f0, "f1": f1, "K": Kmat, "e3": e3, "e2": e2, "gap": o["row_gap"], "idx": o["row_idx"],
           "pairs": set(zip(i0.tolist(), i1.tolist())), "x0": x0, "x1": x1}
    del o
    _CACHE[key] = out
    return out


def run_kind(mv, bb, models, kind, dtype):
    C_ = mv.correspondence
    thr3, thr2 = (NAVI_THR3, NAVI_THR2) if kind == "navi" else (SCAN_THR3, SCAN_THR2)
    hits_g = torch.zeros(len(thr3) + len(thr2))
    hits_r = torch.zeros(len(thr3) + len(thr2))
    total, worst_pair, fracs, common = 0, 0.0, [], []
    C_.set_match_precision(dtype=dtype)
    try:
        for seed in SEEDS:
            o = oracle_case(kind, seed, bb, models)
            p = o["p"]
            if kind == "navi":
                got = C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
            else:
                got = C_.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 1000)
            _, i = C_.knn_points(o["f0"], o["f1"], 2, "cosine")
            clear = o["gap"] > 1e-3
            fracs.append(float(clear.float().mean()))
            if dtype != "bf16" or kind == "navi":
                bad = int((i[clear, 0] != o["idx"][clear, 0]).sum())
                assert bad == 0, f"{kind} seed {seed} {dtype}: {bad} NN mismatches on rows with gap > 1e-3"
            e3, e2 = restated.pair_errors(got[0], got[1], p["Rt"], o["K"])
            g = torch.tensor([(e3 < t).sum().item() for t in thr3] + [(e2 < t).sum().item() for t in thr2], dtype=torch.float32)
            r = torch.tensor([(o["e3"] < t).sum().item() for t in thr3] + [(o["e2"] < t).sum().item() for t in thr2], dtype=torch.float32)
            k = e3.numel()
            assert k == o["e3"].numel() == 1000
            worst_pair = max(worst_pair, float((g - r).abs().max()) * 100.0 / k)
            hits_g += g
            hits_r += r
            total += k
            # how many of the reference's 1000 selected matches we select too (xyz rows identify a match)
            ref_xyz = {tuple(a.tolist()) + tuple(b.tolist()) for a, b in zip(o["x0"][[q[0] for q in o["pairs"]]], o["x1"][[q[1] for q in o["pairs"]]])}
            got_xyz = {tuple(a.tolist()) + tuple(b.tolist()) for a, b in zip(got[0].cpu(), got[1].cpu())}
            common.append(len(ref_xyz & got_xyz))
    finally:
        C_.set_match_precision(dtype=C_.DEFAULT_DTYPE)
    agg = float(((hits_g - hits_r).abs() * 100.0 / total).max())
    return {"agg_pp": agg, "worst_pair_pp": worst_pair, "gap_frac": fracs, "common": common,
            "recall_ref": (hits_r * 100.0 / total).tolist(), "recall_got": (hits_g * 100.0 / total).tolist()}


@pytest.mark.parametrize("dtype", ["f16", "tf32", "bf16"])
def test_navi_vit_features(mv, bb, models, dtype):
    r = run_kind(mv, bb, models, "navi", dtype)
    print(f"NAVI-shaped, random-init ViT-B/16 4-block concat, {dtype}: {r}")
    assert r["agg_pp"] <= 0.1 + 1e-6 and r["worst_pair_pp"] <= 0.2 + 1e-6, r
    assert min(r["gap_frac"]) > 0.80, r          # DESIGN.md section 4 publishes 86-88 %
    assert 20.0 < r["recall_ref"][0] < 95.0, r   # the workload discriminates: recall is neither 0 nor 100
    assert min(r["common"]) >= 995, r


def test_navi_vit_features_lowrank_forced(mv, bb, models):
    """the low-rank proposal (kernel 2 over the 784 source pixels of the target, bicubic taps) is not the default at the
    NAVI shape (no faster there), but it has to hold the same gates when switched on."""
    mv.correspondence.set_match_precision(lowrank=1)
    try:
        r = run_kind(mv, bb, models, "navi", "f16")
    finally:
        mv.correspondence.set_match_precision(lowrank="auto")
    print(f"NAVI-shaped, ViT features, low-rank proposal forced: {r}")
    assert r["agg_pp"] <= 0.1 + 1e-6 and r["worst_pair_pp"] <= 0.2 + 1e-6, r
    assert min(r["common"]) >= 995, r


def test_scannet_resnet_features_dense_product(mv, bb, models):
    """the ScanNet-shaped pairs with the low-rank proposal switched OFF (kernel 2 over the 2056 f16c columns, round 2's
    first form): the same gates."""
    mv.correspondence.set_match_precision(lowrank=0)
    try:
        r = run_kind(mv, bb, models, "scannet", "f16")
    finally:
        mv.correspondence.set_match_precision(lowrank="auto")
    print(f"ScanNet-shaped, ResNet features, dense f16c product: {r}")
    assert r["agg_pp"] <= 0.1 + 1e-6 and r["worst_pair_pp"] <= 0.2 + 1e-6, r
    assert min(r["common"]) >= 990, r


def test_scannet_resnet_features_lowrank_with_row_planes(mv, bb, models):
    """the low-rank proposal with kernel 3 on kernel 1's row planes (lowrank_k3=0) instead of the exact Gram matrix"""
    mv.correspondence.set_match_precision(lowrank_k3=0)
    try:
        r = run_kind(mv, bb, models, "scannet", "f16")
    finally:
        mv.correspondence.set_match_precision(lowrank_k3=1)
    print(f"ScanNet-shaped, ResNet features, low-rank proposal + row planes: {r}")
    assert r["agg_pp"] <= 0.1 + 1e-6 and r["worst_pair_pp"] <= 0.2 + 1e-6, r
    assert min(r["common"]) >= 990, r


@pytest.mark.parametrize("dtype", ["f16", "tf32"])
def test_scannet_resnet_features(mv, bb, models, dtype):
    """the centred operand forms (f16c, the default, and tf32c) on all-positive, nearly collinear CNN features."""
    r = run_kind(mv, bb, models, "scannet", dtype)
    print(f"ScanNet-shaped, random-init ResNet-50 layer4, {dtype}: {r}")
    assert r["agg_pp"] <= 0.1 + 1e-6 and r["worst_pair_pp"] <= 0.2 + 1e-6, r
    assert 20.0 < r["recall_ref"][0] < 95.0, r
    assert min(r["common"]) >= 990, r
    # only ~0.2-0.5 % of the rows have a gap above 1e-3 on these features: the index rule alone says little here,
    # which is why the selected-match overlap and the recall are asserted as well
    assert max(r["gap_frac"]) < 0.05, r


def test_scannet_resnet_features_plain_bf16_is_outside_the_gate(mv, bb, models):
    """documents why plain bf16 rows are not the default operand type: on these features their product proposes the
    wrong neighbours (~ -9 pp recall, half of the selected matches differ; a plain, uncentred tf32 product measured
    0.35 pp / 950-984 matches in common, which is why the tf32 type is centred too).  The centred fp16 form has to be
    strictly better on the same inputs and inside the gate."""
    f = run_kind(mv, bb, models, "scannet", "f16")
    b = run_kind(mv, bb, models, "scannet", "bf16")
    print(f"ScanNet-shaped ResNet features, aggregate recall difference / matches in common with the reference:\n"
          f"  f16c {f['agg_pp']:.2f} pp {f['common']}\n  bf16 {b['agg_pp']:.2f} pp {b['common']}")
    assert min(f["common"]) > max(b["common"])
    assert f["agg_pp"] <= 0.1 + 1e-6 and b["agg_pp"] > 0.1


@pytest.mark.parametrize("dtype", ["f16", "bf16", "tf32"])
def test_spair_vit_features(mv, bb, models, dtype):
    """SPair-shaped pairs from the random-init ViT-B/16 @ 224 through the per-pair path (kernels 1-3) at every operand
    type and through the batched fp32 kernel: arg-max identical to the reference wherever its heat-map top-2 gap exceeds
    1e-3 (north-star rule; the batch kernel is held to 1e-5), key points and errors equal."""
    sp, C_ = mv.spair, mv.correspondence
    pairs = [bb.spair_backbone_pair(s, models["vit_last"], device="cuda", noise=0.5) for s in range(8)]
    es, en, inn, pred = sp.compute_errors_batch(torch.stack([q["feats"] for q in pairs]), torch.stack([q["kps_i"] for q in pairs]),
                                                torch.stack([q["kps_j"] for q in pairs]), [q["thresh_scale"] for q in pairs], 224)
    compared = 0
    C_.set_match_precision(dtype=dtype)
    try:
        for b, q in enumerate(pairs):
            oes, oen, oisame, oinn, heat = restated.spair_compute_errors(q["feats"], q["kps_i"], q["kps_j"], q["thresh_scale"], 224, return_pred=True)
            flat = heat.flatten(1)
            top2 = torch.topk(flat, 2, dim=1).values
            gap = top2[:, 0] - top2[:, 1]
            ges, gen, gisame, ginn, gpred = sp.compute_errors_from_features(q["feats"], q["kps_i"], q["kps_j"], q["thresh_scale"], 224,
                                                                            return_heatmap_argmax=True)
            ok = gap > 1e-3
            assert torch.equal(gpred[ok], flat.argmax(1)[ok])
            assert torch.equal(gisame, oisame)
            sure = ok[oisame]
            torch.testing.assert_close(ges[sure], oes[sure], rtol=1e-5, atol=1e-5)
            okb = gap > 1e-5
            assert torch.equal(pred[b].cpu().long()[okb], flat.argmax(1)[okb])
            compared += int(ok.sum())
    finally:
        C_.set_match_precision(dtype=C_.DEFAULT_DTYPE)
    assert compared >= 8 * 20 * 0.5, compared  # at least half of all key points have a decisive heat map
