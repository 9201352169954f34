"""End-to-end parity of the reference-facing helpers (called with HOST tensors, like the NAVI / ScanNet
callers do) against the golden vectors of the reference and against the oracle at BASELINE.json sizes.

Tolerances (north star): NN indices identical wherever the oracle's fp32 top-2 similarity gap exceeds 1e-3;
recall within 0.1 percentage points; ratio weights to 2e-3 absolute where the neighbours agree."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restated

DEFAULT_DTYPE = "f16"  # correspondence.DEFAULT_DTYPE: what the finally blocks restore

pytestmark = pytest.mark.gpu

SCANNET_SMALL = dict(C=64, h=6, w=8, H=24, W=32)
NAVI_SMALL = dict(C=64, h=8, w=8, H=32, W=32, radius=12.0)
THR3 = [0.01, 0.02, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5]
THR2 = [1, 2, 5, 15, 25, 35, 50]


def t(a):
    return torch.from_numpy(np.asarray(a))


def set_equal_modulo_ties(w_got, w_ref, tol):
    """top-k by weight: the sorted weight vectors must agree to tol (the selected sets can only differ among
    candidates whose weights are within tol of each other)."""
    assert w_got.shape == w_ref.shape
    assert (w_got - w_ref).abs().max() <= tol, float((w_got - w_ref).abs().max())


@pytest.mark.parametrize("dtype", ["f16", "tf32", "bf16"])
@pytest.mark.parametrize("tag,coherent", [("coh", True), ("rnd", False)])
def test_depth_helper_vs_reference_golden(mv, syn, golden, dtype, tag, coherent):
    g = golden(f"scannet_small_{tag}")
    p = syn.scannet_pair(7, coherent=coherent, **SCANNET_SMALL)
    mv.correspondence.set_match_precision(dtype=dtype)
    try:
        x0, x1, w = mv.correspondence.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 100)
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    assert x0.device.type == "cpu" and x0.shape == (100, 3) and w.shape == (100,)
    assert (w[1:] <= w[:-1]).all()
    tol = 2e-2 if dtype == "bf16" else 2e-3  # the default (f16c) is held to the tf32 tolerance
    set_equal_modulo_ties(w, t(g["corr_dist"]), tol)
    # every returned match whose weight is clear of the k-th weight by tol must be in the reference's set
    ref0 = {tuple(np.round(r, 5)) for r in g["corr_xyz0"]}
    clear = w > (w[-1] + tol)
    got0 = [tuple(np.round(r, 5)) for r in x0[clear].numpy()]
    assert sum(r in ref0 for r in got0) >= (0.97 if dtype == "bf16" else 0.99) * len(got0)
    e3, _ = restated.pair_errors(x0, x1, p["Rt"], p["K"])
    for th in THR3:
        r_got = 100.0 * (e3 < th).float().mean().item()
        r_ref = 100.0 * float((g["err3d"] < th).mean())
        # k = 100 matches: one match = 1 pp; the 0.1 pp gate is tested at full size
        assert abs(r_got - r_ref) <= (2.0 if dtype == "bf16" else 1.0) + 1e-4


@pytest.mark.parametrize("tag,coherent", [("coh", True), ("rnd", False)])
def test_xyz_helper_vs_reference_golden(mv, syn, golden, tag, coherent):
    g = golden(f"navi_small_{tag}")
    p = syn.navi_pair(7, coherent=coherent, **NAVI_SMALL)
    mv.correspondence.set_match_precision(dtype="tf32")
    try:
        out = mv.correspondence.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 100)
        nr = mv.correspondence.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 100, ratio_test=False)
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    assert [tuple(o.shape) for o in out] == [(100, 3), (100, 3), (100,), (100, 2), (100, 2)]
    set_equal_modulo_ties(out[2], t(g["c_dist"]), 2e-3)
    set_equal_modulo_ties(nr[2], t(g["nr_dist"]), 1e-5)
    # uv are pixel centres of the gathered grid positions
    assert ((out[3] - 0.5) == (out[3] - 0.5).round()).all()


def test_rows_helper_vs_reference_golden(mv, golden):
    g = golden("rows_small")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    X = torch.randn(300, 64, generator=gen)
    Y = torch.randn(280, 64, generator=gen)
    C_ = mv.correspondence
    C_.set_match_precision(dtype="tf32")
    try:
        d, i = C_.knn_points(X, Y, 2, "cosine")
        i1, i2, w = C_.get_correspondences_ratio_test(X, Y, 50)
        b1, b2, bw = C_.get_correspondences_ratio_test(X, Y, 50, bidirectional=True)
    finally:
        C_.set_match_precision(dtype=DEFAULT_DTYPE)
    o = restated.similarity_top2_and_mutual(X, Y)
    clear = o["row_gap"] > 1e-3
    assert i.dtype == torch.int64 and torch.equal(i[clear, 0], t(g["idx"])[clear, 0])
    both = torch.equal(i, t(g["idx"]))
    if both:
        torch.testing.assert_close(d, t(g["dists"]), rtol=0, atol=1e-6)
        assert torch.equal(i1, t(g["idx1"])) and torch.equal(i2, t(g["idx2"]))
        torch.testing.assert_close(w, t(g["weight"]), rtol=0, atol=1e-5)
    assert b1.shape == (50,) and bw.shape == (50,)
    s, tg, v = C_.get_topk_matches(t(g["ratio"]), t(g["idx"])[:, 0], 50)
    assert torch.equal(s, t(g["idx1"])) and torch.equal(tg, t(g["idx2"]))
    torch.testing.assert_close(v, t(g["weight"]), rtol=0, atol=0)
    torch.testing.assert_close(C_.calculate_ratio_test(t(g["dists"])), t(g["ratio"]), rtol=0, atol=0)
    torch.testing.assert_close(C_.get_grid(3, 5), t(g["grid"]), rtol=0, atol=0)


def test_faiss_knn_and_euclidean(mv):
    gen = torch.Generator().manual_seed(9)
    X = torch.randn(200, 40, generator=gen) * 3
    Y = torch.randn(150, 40, generator=gen) * 3
    d, i = mv.correspondence.faiss_knn(X, Y, 2)
    od, oi = restated.exact_l2_knn(X, Y, 3)
    # the north-star gap rule on squared distances: neighbours are compared wherever consecutive distances differ by more
    # than 1e-3 relative (the tf32 product over mean-centred targets proposes, fp32 distances of the two decide)
    clear = (od[:, 2] - od[:, 1] > 1e-3 * od[:, 1]) & (od[:, 1] - od[:, 0] > 1e-3 * od[:, 0])
    assert clear.float().mean() >= 0.9
    assert torch.equal(i[clear], oi[clear][:, :2])
    torch.testing.assert_close(d[clear], od[clear][:, :2], rtol=1e-4, atol=1e-3)
    de, ie = mv.correspondence.knn_points(X, Y, 1, "euclidean")
    assert torch.equal(ie[clear, 0], oi[clear, 0])


def test_more_than_two_neighbours(mv):
    """K > 2 (no call site in the reference): exact blocked search on the device, same interface."""
    gen = torch.Generator().manual_seed(33)
    X = torch.randn(500, 48, generator=gen)
    Y = torch.randn(700, 48, generator=gen)
    d, i = mv.correspondence.faiss_knn(X, Y, 5)
    od, oi = restated.exact_l2_knn(X, Y, 5)
    assert i.dtype == torch.int64 and i.shape == (500, 5)
    assert (i == oi).float().mean() > 0.999
    torch.testing.assert_close(d, od, rtol=1e-4, atol=1e-3)
    for metric in ("cosine", "euclidean"):
        dk, ik = mv.correspondence.knn_points(X, Y, 4, metric)
        odk, oik = restated.knn_points(X, Y, 4, metric)
        assert (ik == oik).float().mean() > 0.999
        torch.testing.assert_close(dk, odk, rtol=1e-4, atol=1e-5)


def test_ratio_test_euclidean_metric(mv):
    """metric="euclidean" of get_correspondences_ratio_test (correspondence.py:63-102; no call site in the
    reference, kept for interface parity) against the oracle's literal restatement."""
    gen = torch.Generator().manual_seed(21)
    X = torch.randn(300, 48, generator=gen) * 2
    Y = X[torch.randperm(300, generator=gen)][:260] + 0.3 * torch.randn(260, 48, generator=gen)
    i1, i2, w = mv.correspondence.get_correspondences_ratio_test(X, Y, 50, metric="euclidean")
    od, oi = restated.knn_points(X, Y, 2, "euclidean")
    ow = 1 - od[:, 0].clamp(min=1e-9) / od[:, 1].clamp(min=1e-9)
    wt, it = torch.topk(ow, 50)
    assert i1.dtype == torch.int64 and w.shape == (50,)
    assert set(i1.tolist()) == set(it.tolist())
    torch.testing.assert_close(w, wt, rtol=1e-4, atol=1e-5)
    assert torch.equal(i2, oi[i1, 0])
    j1, j2, jw = mv.correspondence.get_correspondences_ratio_test(X, Y, 40, metric="euclidean", bidirectional=True, ratio_test=False)
    assert j1.shape == (40,) and j2.shape == (40,) and jw.shape == (40,)


def full_size_case(mv, kind, dtype, syn):
    C_ = mv.correspondence
    if kind == "scannet":
        p = syn.scannet_pair(1, coherent=True)
        _, f0, _ = restated.depth_side(p["feat_0"], p["depth_0"], p["K"])
        _, f1, _ = restated.depth_side(p["feat_1"], p["depth_1"], p["K"])
        Kmat = p["K"]
    else:
        p = syn.navi_pair(1, coherent=True)
        _, f0, _, _ = restated.xyz_side(p["feat_0"], p["xyz_grid_0"])
        _, f1, _, _ = restated.xyz_side(p["feat_1"], p["xyz_grid_1"])
        Kmat = p["intrinsics"]
    C_.set_match_precision(dtype=dtype)
    try:
        if kind == "scannet":
            got = C_.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 1000)
            ref = restated.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 1000)
        else:
            got = C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
            ref = restated.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
        d, i = C_.knn_points(f0, f1, 2, "cosine")
    finally:
        C_.set_match_precision(dtype=DEFAULT_DTYPE)
    o = restated.similarity_top2_and_mutual(f0, f1)
    clear = o["row_gap"] > 1e-3
    frac = float(clear.float().mean())
    # NN indices identical wherever the reference's top-2 similarity gap exceeds 1e-3
    assert torch.equal(i[clear, 0], o["row_idx"][clear, 0]), f"{(i[clear,0] != o['row_idx'][clear,0]).sum()} mismatches"
    e3g, e2g = restated.pair_errors(got[0], got[1], p["Rt"], Kmat)
    e3r, e2r = restated.pair_errors(ref[0], ref[1], p["Rt"], Kmat)
    rec = {}
    for th in THR3:
        a, b = 100.0 * (e3g < th).float().mean().item(), 100.0 * (e3r < th).float().mean().item()
        rec[th] = (a, b)
        assert abs(a - b) <= 0.1 + 1e-6, (kind, dtype, th, a, b)
    for th in THR2:
        a, b = 100.0 * (e2g < th).float().mean().item(), 100.0 * (e2r < th).float().mean().item()
        assert abs(a - b) <= 0.1 + 1e-6, (kind, dtype, th, a, b)
    return frac, rec


@pytest.mark.parametrize("dtype", ["f16", "bf16", "tf32"])
def test_scannet_full_size_parity(mv, syn, dtype):
    frac, rec = full_size_case(mv, "scannet", dtype, syn)
    print(f"scannet-shaped {dtype}: rows compared (gap > 1e-3) = {100 * frac:.1f}%, recall@thr (got, ref) = {rec}")
    assert frac > 0.55  # DESIGN.md publishes 59 % for the Gaussian ScanNet-shaped pair


@pytest.mark.parametrize("dtype", ["f16", "bf16", "tf32"])
def test_navi_full_size_parity(mv, syn, dtype):
    frac, rec = full_size_case(mv, "navi", dtype, syn)
    print(f"navi-shaped {dtype}: rows compared (gap > 1e-3) = {100 * frac:.1f}%, recall@thr (got, ref) = {rec}")
    assert frac > 0.70  # DESIGN.md publishes 73 % for the Gaussian NAVI-shaped pair


def test_fused_scoring_equals_helper_plus_oracle_errors(mv, syn):
    """evaluation.match_and_score_depth (no host sync, device-resident counts) must give the same integer
    counts as the synced helper followed by the oracle's error computation."""
    ev = mv.evaluation
    p = syn.scannet_pair(2, coherent=True, C=256, h=15, w=20, H=60, W=80)
    acc_a = ev.RecallAccumulator(THR3, THR2, device="cuda")
    acc_b = ev.RecallAccumulator(THR3, THR2, device="cuda")
    ra = ev.match_and_score_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"], p["Rt"], 500, acc_a, sync=False)
    rb = ev.match_and_score_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"], p["Rt"], 500, acc_b, sync=True)
    assert acc_a.hits.cpu().tolist() == acc_b.hits.cpu().tolist()
    x0, x1, w = mv.correspondence.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], p["K"].clone(), 500)
    e3, e2 = restated.pair_errors(x0, x1, p["Rt"], p["K"])
    s = acc_a.summary()
    for th in THR3:
        assert abs(s["recall_3d"][th] - 100.0 * (e3 < th).float().mean().item()) <= 0.2 + 1e-6
    assert s["scored"] == 500


@pytest.mark.parametrize("kind", ["xyz", "depth"])
def test_cuda_graph_replay_equals_eager(mv, syn, kind):
    """GraphedPairMatcher (captured once, replayed per pair with inputs staged into static buffers) must give
    the same integer counts and the same selected matches as the eager launches, pair after pair."""
    ev = mv.evaluation
    if kind == "xyz":
        pairs = [syn.navi_pair(i, C=256, h=14, w=14, H=56, W=56, radius=20.0) for i in range(3)]
        gk = ("xyz_grid_0", "xyz_grid_1", "intrinsics")
    else:
        pairs = [syn.scannet_pair(i, C=256, h=15, w=20, H=60, W=80) for i in range(3)]
        gk = ("depth_0", "depth_1", "K")
    gm = ev.GraphedPairMatcher(kind, tuple(pairs[0]["feat_0"].shape), tuple(pairs[0][gk[0]].shape), 300, K=pairs[0].get("K")).capture()
    for p in pairs + pairs[:1]:
        a = ev.RecallAccumulator(THR3, THR2, device="cuda")
        b = ev.RecallAccumulator(THR3, THR2, device="cuda")
        gm.load(p["feat_0"], p["feat_1"], p[gk[0]], p[gk[1]])
        rg = gm.run(a, p["Rt"], p[gk[2]])
        if kind == "xyz":
            re_ = ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p[gk[0]], p[gk[1]], p[gk[2]], p["Rt"], 300, b, sync=True)
        else:
            re_ = ev.match_and_score_depth(p["feat_0"], p["feat_1"], p[gk[0]], p[gk[1]], p[gk[2]], p["Rt"], 300, b, sync=True)
        assert a.hits.cpu().tolist() == b.hits.cpu().tolist()
        k = re_.k
        assert int(rg.k_dev.item()) == k
        assert torch.equal(rg.sel_src[:k], re_.sel_src[:k]) and torch.equal(rg.sel_weight[:k], re_.sel_weight[:k])


def test_pair_pipeline_counts_equal_sequential(mv, syn):
    """Two pairs in flight on separate streams must accumulate exactly the counts of the sequential run."""
    ev = mv.evaluation
    pairs = [syn.navi_pair(i, C=256, h=14, w=14, H=56, W=56, radius=20.0) for i in range(5)]
    shape_f, shape_g = tuple(pairs[0]["feat_0"].shape), tuple(pairs[0]["xyz_grid_0"].shape)
    seq = ev.RecallAccumulator(THR3, THR2, device="cuda")
    for p in pairs:
        ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], p["intrinsics"], p["Rt"], 300, seq, sync=True)
    for lanes in (1, 2, 3):
        acc = ev.RecallAccumulator(THR3, THR2, device="cuda")
        pipe = ev.PairPipeline("xyz", shape_f, shape_g, 300, lanes=lanes)
        for p in pairs:
            pipe.submit(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], acc, p["Rt"], p["intrinsics"])
        pipe.join()
        assert acc.hits.cpu().tolist() == seq.hits.cpu().tolist()


def test_edge_cases_of_the_helpers(mv, syn):
    C_ = mv.correspondence
    g = torch.Generator().manual_seed(0)
    # fewer points than requested correspondences: k = N (correspondence.py:126)
    X, Y = torch.randn(7, 16, generator=g), torch.randn(40, 16, generator=g)
    i1, i2, w = C_.get_correspondences_ratio_test(X, Y, 50)
    assert i1.shape == (7,) and sorted(i1.tolist()) == list(range(7)) and (w[1:] <= w[:-1]).all()
    o1, o2, ow = restated.correspondences_ratio_test(X, Y, 50)
    assert torch.equal(i1, o1) and torch.equal(i2, o2)
    torch.testing.assert_close(w, ow, rtol=0, atol=1e-5)
    # CUDA tensors in -> CUDA tensors out (the SPair caller keeps everything on the device)
    d, i = C_.knn_points(X.cuda(), Y.cuda(), 2, "cosine")
    assert d.is_cuda and i.is_cuda and i.dtype == torch.int64
    # a pair without valid geometry raises instead of returning garbage
    p = syn.navi_pair(0, C=64, h=8, w=8, H=32, W=32, radius=12.0)
    with pytest.raises(RuntimeError, match="too few valid points"):
        C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], torch.zeros_like(p["xyz_grid_0"]), p["xyz_grid_1"], 10)
    with pytest.raises(RuntimeError, match="too few valid points"):
        C_.estimate_correspondence_xyz(p["feat_0"].cuda(), p["feat_1"].cuda(), torch.zeros_like(p["xyz_grid_0"]).cuda(), p["xyz_grid_1"].cuda(), 10)
    # ragged: every other pixel valid, odd sizes, k larger than the live count
    grid = p["xyz_grid_0"].clone()
    grid[2, ::2, :] = 0
    out = C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], grid, p["xyz_grid_1"], 100000)
    ref = restated.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], grid, p["xyz_grid_1"], 100000)
    assert out[0].shape == ref[0].shape
    assert (out[2] - ref[2]).abs().max() < 2e-2
    with pytest.raises(ValueError):
        C_.estimate_correspondence_xyz(torch.zeros(7, 4, 4), torch.zeros(7, 4, 4), p["xyz_grid_0"], p["xyz_grid_1"], 10)


def test_channel_last_features_take_the_zero_copy_path(mv, syn):
    """a (C, h, w) view of channel-last memory (ViT tokens before tokens_to_output's .contiguous()) gives the same
    results as the reference's contiguous layout, through the eager helper, the cached-graph helper and the
    pipeline's static buffers."""
    C_, ev = mv.correspondence, mv.evaluation
    p = syn.navi_pair(11, C=256, h=12, w=12, H=48, W=48, radius=18.0)
    cl = lambda t: t.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    assert C_._is_channel_last(cl(p["feat_0"])) and not C_._is_channel_last(p["feat_0"])
    for graphs in (0, 1):
        C_.set_match_precision(helper_graphs=graphs)
        try:
            a = C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 200)
            b = C_.estimate_correspondence_xyz(cl(p["feat_0"]), cl(p["feat_1"]), p["xyz_grid_0"], p["xyz_grid_1"], 200)
        finally:
            C_.set_match_precision(helper_graphs=1)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    d = syn.scannet_pair(12, C=64, h=6, w=8, H=24, W=32)
    a = C_.estimate_correspondence_depth(d["feat_0"].cuda(), d["feat_1"].cuda(), d["depth_0"].cuda(), d["depth_1"].cuda(), d["K"], 100)
    b = C_.estimate_correspondence_depth(cl(d["feat_0"]).cuda(), cl(d["feat_1"]).cuda(), d["depth_0"].cuda(), d["depth_1"].cuda(), d["K"], 100)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    accs = []
    for layout in ("chw", "hwc"):
        acc = ev.RecallAccumulator([0.01, 0.02], [5, 25], device=torch.device("cuda"))
        gm = ev.GraphedPairMatcher("xyz", tuple(p["feat_0"].shape), tuple(p["xyz_grid_0"].shape), 200, feat_layout=layout).capture()
        f0, f1 = (cl(p["feat_0"]), cl(p["feat_1"])) if layout == "hwc" else (p["feat_0"], p["feat_1"])
        gm.load(f0.cuda(), f1.cuda(), p["xyz_grid_0"].cuda(), p["xyz_grid_1"].cuda())
        gm.run(acc, p["Rt"], p["intrinsics"])
        accs.append(acc.hits.cpu())
    assert torch.equal(accs[0], accs[1]) and int(accs[0][0]) == 200


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_16bit_feature_maps_equal_their_fp32_widening(mv, syn, dt):
    """autocast backbones hand over bf16 / fp16 maps: they are uploaded as they are (half the bytes) and widened on
    the device, which is exact -- the outputs must be bit-identical to the call with feat.float(), in both memory
    layouts, through the eager helper, the cached-graph helper and the 16-bit static buffers of the graph matcher."""
    C_, ev = mv.correspondence, mv.evaluation
    p = syn.navi_pair(13, C=256, h=12, w=12, H=48, W=48, radius=18.0)
    f0, f1 = p["feat_0"].to(dt), p["feat_1"].to(dt)
    cl = lambda t: t.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    for graphs in (0, 1):
        C_.set_match_precision(helper_graphs=graphs)
        try:
            want = C_.estimate_correspondence_xyz(f0.float(), f1.float(), p["xyz_grid_0"], p["xyz_grid_1"], 200)
            for a, b in ((f0, f1), (cl(f0), cl(f1))):
                got = C_.estimate_correspondence_xyz(a, b, p["xyz_grid_0"], p["xyz_grid_1"], 200)
                for x, y in zip(want, got):
                    assert torch.equal(x, y)
        finally:
            C_.set_match_precision(helper_graphs=1)
    d = syn.scannet_pair(14, C=64, h=6, w=8, H=24, W=32)
    g0, g1 = d["feat_0"].to(dt).cuda(), d["feat_1"].to(dt).cuda()
    want = C_.estimate_correspondence_depth(g0.float(), g1.float(), d["depth_0"].cuda(), d["depth_1"].cuda(), d["K"], 100)
    got = C_.estimate_correspondence_depth(g0, g1, d["depth_0"].cuda(), d["depth_1"].cuda(), d["K"], 100)
    for x, y in zip(want, got):
        assert torch.equal(x, y)
    accs = []
    for fdt, a, b in ((torch.float32, f0.float(), f1.float()), (dt, f0, f1)):
        acc = ev.RecallAccumulator([0.01, 0.02], [5, 25], device=torch.device("cuda"))
        gm = ev.GraphedPairMatcher("xyz", tuple(f0.shape), tuple(p["xyz_grid_0"].shape), 200, feat_dtype=fdt).capture()
        gm.load(a.cuda(), b.cuda(), p["xyz_grid_0"].cuda(), p["xyz_grid_1"].cuda())
        gm.run(acc, p["Rt"], p["intrinsics"])
        accs.append(acc.hits.cpu())
    assert torch.equal(accs[0], accs[1]) and int(accs[0][0]) == 200


def test_second_device_in_the_same_process(mv, syn):
    """function attributes (large shared-memory opt-ins, cluster occupancy) are cached per device: a process that
    uses cuda:0 and then cuda:1 must get the same results on both."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    C_ = mv.correspondence
    p = syn.navi_pair(15, C=3072, h=12, w=12, H=48, W=48, radius=18.0)
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            outs.append([t.cpu() for t in C_.estimate_correspondence_xyz(p["feat_0"].cuda(d), p["feat_1"].cuda(d), p["xyz_grid_0"].cuda(d),
                                                                        p["xyz_grid_1"].cuda(d), 300)])
            s = syn.spair_pair(3)
            e = mv.spair.compute_errors_batch(s["feats"][None].cuda(d), s["kps_i"][None], s["kps_j"][None], [s["thresh_scale"]], 224)
            outs[-1].append(e[3].cpu())
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_intrinsics_change_between_replays_of_one_graph(mv, syn):
    """ScanNet's intrinsics differ per scene: the cached graph keeps K / K^-1 in device memory, so calls with
    different K re-use ONE captured graph and still equal the eager launches with that K."""
    C_ = mv.correspondence
    p = syn.scannet_pair(16, C=64, h=6, w=8, H=24, W=32)
    Ks = [p["K"].clone(), p["K"].clone() * torch.tensor([[1.05], [0.97], [1.0]]), p["K"].clone()]
    Ks[1][0, 2] += 1.5
    n_before = len(C_._HELPER_GRAPHS)
    outs_g = [C_.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], K.clone(), 100) for K in Ks]
    assert len(C_._HELPER_GRAPHS) <= n_before + 1                      # one graph for all three K
    C_.set_match_precision(helper_graphs=0)
    try:
        outs_e = [C_.estimate_correspondence_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], K.clone(), 100) for K in Ks]
    finally:
        C_.set_match_precision(helper_graphs=1)
    for a, b in zip(outs_g, outs_e):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert not torch.equal(outs_g[0][0], outs_g[1][0]) and torch.equal(outs_g[0][0], outs_g[2][0])
    # the pipeline takes per-pair intrinsics too
    ev = mv.evaluation
    acc_a = ev.RecallAccumulator([0.05], [25], device=torch.device("cuda"))
    acc_b = ev.RecallAccumulator([0.05], [25], device=torch.device("cuda"))
    pipe = ev.PairPipeline("depth", tuple(p["feat_0"].shape), tuple(p["depth_0"].shape), 100, K=Ks[0], lanes=2)
    for K in Ks:
        pipe.submit(p["feat_0"].cuda(), p["feat_1"].cuda(), p["depth_0"].cuda(), p["depth_1"].cuda(), acc_a, p["Rt"], K)
        ev.match_and_score_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], K, p["Rt"], 100, acc_b, sync=True)
    pipe.join()
    torch.cuda.synchronize()
    assert torch.equal(acc_a.hits.cpu(), acc_b.hits.cpu())


def test_many_intrinsics_changes_queued_behind_a_long_kernel(mv, syn):
    """the matcher stages [K | K^-1] in a ring of 8 pinned rows; with the host running far ahead of the GPU (a long
    kernel in front, 20 K changes queued behind it on ONE lane) a staging row must not be rewritten before the copy
    that reads it has run -- every pair has to be scored with ITS intrinsics."""
    ev = mv.evaluation
    p = syn.scannet_pair(17, C=64, h=6, w=8, H=24, W=32)
    dev = torch.device("cuda")
    T3, T2 = [0.004, 0.006, 0.008, 0.012, 0.02], [1.5, 3.0]  # around the 3-D distance of neighbouring pixels: sensitive to K
    Ks = []
    for i in range(20):
        K = p["K"].clone()
        K[0, 0] *= 1.0 + 0.03 * i
        K[1, 1] *= 1.0 - 0.02 * i
        K[1, 2] += 0.3 * i
        Ks.append(K)
    want = []
    for K in Ks:
        acc = ev.RecallAccumulator(T3, T2, device=dev)
        ev.match_and_score_depth(p["feat_0"], p["feat_1"], p["depth_0"], p["depth_1"], K, p["Rt"], 100, acc, sync=True)
        want.append(acc.hits.cpu())
    assert len({tuple(w.tolist()) for w in want}) > 3, want  # the intrinsics matter for the counts
    pipe = ev.PairPipeline("depth", tuple(p["feat_0"].shape), tuple(p["depth_0"].shape), 100, K=Ks[0], lanes=1)
    f0, f1, d0, d1 = (p[k].cuda() for k in ("feat_0", "feat_1", "depth_0", "depth_1"))
    accs = [ev.RecallAccumulator(T3, T2, device=dev) for _ in Ks]
    st, _ = pipe.lanes[0]
    with torch.cuda.stream(st):
        torch.cuda._sleep(int(2e8))  # ~0.1 s of GPU time in front of everything that follows on the lane
    for K, acc in zip(Ks, accs):
        pipe.submit(f0, f1, d0, d1, acc, p["Rt"], K)
    pipe.join()
    torch.cuda.synchronize()
    for a, w in zip(accs, want):
        assert torch.equal(a.hits.cpu(), w)


def test_differently_sized_images_take_the_eager_path(mv, syn):
    """the reference accepts two images of different sizes; the cached-graph fast path is keyed on image 0's shapes, so
    such a call must fall back to the eager launches instead of copying image 1 into image 0's buffers."""
    C_ = mv.correspondence
    a = syn.navi_pair(21, C=64, h=8, w=8, H=32, W=32, radius=12.0)
    b = syn.navi_pair(22, C=64, h=6, w=6, H=24, W=24, radius=9.0)
    got = C_.estimate_correspondence_xyz(a["feat_0"], b["feat_1"], a["xyz_grid_0"], b["xyz_grid_1"], 50)
    ref = restated.estimate_correspondence_xyz(a["feat_0"], b["feat_1"], a["xyz_grid_0"], b["xyz_grid_1"], 50)
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in ref]
    torch.testing.assert_close(got[2], ref[2], rtol=0, atol=2e-3)


def test_bidirectional_ratio_test_vs_oracle(mv):
    """bidirectional=True (correspondence.py:79-98 with its concatenations along dim 0; the reference's own branch
    raises): half the budget per direction, against the oracle's restatement of those lines."""
    C_ = mv.correspondence
    gen = torch.Generator().manual_seed(44)
    X = torch.randn(400, 64, generator=gen)
    Y = X[torch.randperm(400, generator=gen)][:350] + 0.5 * torch.randn(350, 64, generator=gen)
    for ratio in (True, False):
        i1, i2, w = C_.get_correspondences_ratio_test(X, Y, 60, bidirectional=True, ratio_test=ratio)
        o1, o2, ow = restated.correspondences_ratio_test_bidirectional(X, Y, 60, ratio_test=ratio)
        assert i1.dtype == torch.int64 and i1.shape == (60,)
        torch.testing.assert_close(w, ow, rtol=0, atol=2e-5)
        clear = torch.ones(60, dtype=torch.bool)
        for half in (slice(0, 30), slice(30, 60)):  # inside each direction: exact unless two weights are within 2e-5
            gaps = (ow[half][:-1] - ow[half][1:]).abs()
            ok = torch.ones(30, dtype=torch.bool)
            ok[:-1] &= gaps > 4e-5
            ok[1:] &= gaps > 4e-5
            ok[-1] = False  # the last of a half competes with the first one left out
            clear[half] = ok
        assert clear.float().mean() > 0.8
        assert torch.equal(i1[clear], o1[clear]) and torch.equal(i2[clear], o2[clear])


def test_reference_caller_runs_through_the_module_alias(mv, syn):
    """INTEGRATION.md option B: alias this package's modules as evals.utils.correspondence / evals.utils.transformations
    and run a reference CALLER on top -- the NAVI evaluation loop (evaluate_navi_correspondence.py:174-221; its
    restatement oracle/restated.navi_error_block is pinned against the reference's own text by tests/test_oracle.py),
    which imports estimate_correspondence_xyz / project_3dto2d / compute_binned_performance / transform_points_Rt /
    so3_rotation_angle by name and feeds CPU tensors, exactly like the script.  Same recalls as the same loop over the
    oracle's functions."""
    import importlib
    import sys
    import types

    saved = {k: sys.modules.get(k) for k in ("evals", "evals.utils", "evals.utils.correspondence", "evals.utils.transformations")}
    try:
        pkg, sub = types.ModuleType("evals"), types.ModuleType("evals.utils")
        pkg.utils = sub
        sys.modules["evals"], sys.modules["evals.utils"] = pkg, sub
        sys.modules["evals.utils.correspondence"] = mv.correspondence
        sys.modules["evals.utils.transformations"] = mv.transformations
        corr = importlib.import_module("evals.utils.correspondence")
        tr = importlib.import_module("evals.utils.transformations")
        from evals.utils.correspondence import estimate_correspondence_xyz  # noqa: F401  (the script's own import line works)

        ps = [syn.navi_pair(60 + i, coherent=True, C=256, h=14, w=14, H=56, W=56, radius=20.0) for i in range(4)]
        g = torch.Generator().manual_seed(78)
        Rt = torch.stack([syn.random_rt(g, max_deg=110.0) for _ in ps])
        for p, r in zip(ps, Rt):  # make the geometry consistent with the pair's Rt: image 1 sees the same points moved by Rt
            p["xyz_grid_1"] = (r[:3, :3] @ p["xyz_grid_0"].reshape(3, -1) + r[:3, 3:4]).reshape(p["xyz_grid_0"].shape) * (p["xyz_grid_0"][2:3] > 0)
            p["xyz_grid_1"][2] = torch.where(p["xyz_grid_0"][2] > 0, p["xyz_grid_1"][2].clamp(min=1e-3), torch.zeros(()))
        stack = lambda k: torch.stack([p[k] for p in ps])
        args = (stack("feat_0"), stack("feat_1"), stack("xyz_grid_0"), stack("xyz_grid_1"), Rt, stack("intrinsics"))
        got = restated.navi_error_block(corr, tr, *args, num_corr=300)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    oracle_mod = types.SimpleNamespace(estimate_correspondence_xyz=restated.estimate_correspondence_xyz, project_3dto2d=restated.project_3dto2d,
                                       compute_binned_performance=restated.compute_binned_performance)
    oracle_tr = types.SimpleNamespace(transform_points_Rt=restated.transform_points_Rt, so3_rotation_angle=restated.so3_rotation_angle)
    want = restated.navi_error_block(oracle_mod, oracle_tr, *args, num_corr=300)
    assert got["err_3d"].shape == want["err_3d"].shape == (4, 300) and got["err_3d"].device.type == "cpu"
    # 300 matches per pair: one match = 0.33 pp per pair, 0.083 pp on the aggregate
    assert (got["recall_3d"] - want["recall_3d"]).abs().max() <= 0.1 + 1e-4
    assert (got["recall_2d"] - want["recall_2d"]).abs().max() <= 0.1 + 1e-4
    gb, wb = got["bin_rec"], want["bin_rec"]
    assert torch.equal(gb.isnan(), wb.isnan()) and (gb[~gb.isnan()] - wb[~wb.isnan()]).abs().max() <= 0.004
    assert 5.0 < float(want["recall_3d"][0]) < 100.0


def test_in_process_sharding_gives_the_single_run_counts(mv, syn):
    """pair i -> shard i mod W, one accumulator per shard, summed at the end: the same integers as one accumulator over
    all pairs -- real kernels, one GPU (the NCCL form of the same statement is test_two_gpu_nccl_counts below and
    bench.py's sharded_set digest)."""
    ev = mv.evaluation
    dev = torch.device("cuda")
    pairs = [syn.navi_pair(80 + i, C=256, h=14, w=14, H=56, W=56, radius=14.0 + i) for i in range(9)]
    T3, T2 = [0.01, 0.02, 0.05], [5, 25, 50]

    def run(idx, acc):
        for i in idx:
            p = pairs[i]
            ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], p["intrinsics"], p["Rt"], 200, acc, sync=True)

    single = ev.RecallAccumulator(T3, T2, device=dev)
    run(range(9), single)
    for W in (2, 4):
        total = torch.zeros_like(single.hits)
        for r in range(W):
            acc = ev.RecallAccumulator(T3, T2, device=dev)
            run(ev.shard_pairs(9, r, W), acc)
            total += acc.hits
        assert torch.equal(total, single.hits)


def _nccl_worker(rank, world, port, out_dir):
    import importlib
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    mv_ = importlib.import_module("midvision-probe_b200")
    syn_ = importlib.import_module("midvision-probe_b200.synthetic")
    ev = mv_.evaluation
    acc = ev.RecallAccumulator([0.01, 0.02, 0.05], [5, 25, 50], device=dev)
    for i in ev.shard_pairs(9, rank, world):
        p = syn_.navi_pair(80 + i, C=256, h=14, w=14, H=56, W=56, radius=14.0 + i)
        ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], p["intrinsics"], p["Rt"], 200, acc, sync=True)
    acc.all_reduce()
    torch.save(acc.hits.cpu(), os.path.join(out_dir, f"nccl_r{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_gpu_nccl_counts(tmp_path, mv, syn):
    """two ranks, two GPUs, NCCL, real kernels: after the single all-reduce both ranks hold the integers of the
    single-GPU run.  Skipped on a one-GPU box (the driver's); run with `gpurun --gpus 2`."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os

    import torch.multiprocessing as mp

    ev = mv.evaluation
    single = ev.RecallAccumulator([0.01, 0.02, 0.05], [5, 25, 50], device=torch.device("cuda", 0))
    for i in range(9):
        p = syn.navi_pair(80 + i, C=256, h=14, w=14, H=56, W=56, radius=14.0 + i)
        ev.match_and_score_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], p["intrinsics"], p["Rt"], 200, single, sync=True)
    port = 29700 + os.getpid() % 200
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert torch.equal(torch.load(os.path.join(tmp_path, f"nccl_r{r}.pt")), single.hits.cpu())


def test_staged_upload_of_pageable_memory_is_exact(mv):
    """mv_h2d_staged (threaded pinned ring, csrc/stage.cu): odd sizes, transfers larger than the 16 MiB ring, many calls back
    to back on one stream (the ring position runs on across calls) and behind a long kernel -- the device copy must equal the
    pageable source bit for bit, also when the source is overwritten right after the call returns (every chunk is staged by then)."""
    from ctypes import c_void_p

    L = mv._lib
    st = torch.cuda.current_stream()
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 4096, (1 << 18) - 3, (1 << 18) + 5, 9633792, (1 << 24) + 12345, 40_000_001, 150_528, 9633792, 150_528]
    busy = torch.randn(4096, 4096, device="cuda")
    dsts, refs = [], []
    for rep in range(2):
        for _ in range(3):
            busy = busy @ busy * 1e-3  # keep the stream busy so that ring slots have to wait for their DMAs
        for nb in sizes:
            src = torch.randint(0, 256, (nb,), dtype=torch.uint8, generator=g)
            dst = torch.zeros(nb, dtype=torch.uint8, device="cuda")
            L.call("mv_h2d_staged", c_void_p(dst.data_ptr()), c_void_p(src.data_ptr()), nb, c_void_p(st.cuda_stream))
            refs.append(src.clone())
            src.zero_()  # the call has returned: the source may be modified
            dsts.append(dst)
    torch.cuda.synchronize()
    for d, r in zip(dsts, refs):
        assert torch.equal(d.cpu(), r)
