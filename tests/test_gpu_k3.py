"""Kernel 3 pieces against torch / the oracle: distances, ratio weights, mutual flags, top-k, scoring, SPair."""
from ctypes import c_float, c_void_p

import pytest
import torch
import torch.nn.functional as F

from oracle import restated

DEFAULT_DTYPE = "f16"  # correspondence.DEFAULT_DTYPE: what the finally blocks restore

pytestmark = pytest.mark.gpu


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n,m,C", [(300, 280, 64), (1, 2, 8), (1000, 900, 768), (257, 513, 2048)])
@pytest.mark.parametrize("ratio_test", [True, False])
def test_ratio_mutual(mv, n, m, C, ratio_test):
    L = mv._lib
    g = torch.Generator().manual_seed(n + C)
    X = F.normalize(torch.randn(n, C, generator=g), dim=1)
    Y = F.normalize(torch.randn(m, C, generator=g), dim=1)
    o = restated.similarity_top2_and_mutual(X, Y)
    idx = o["row_idx"].clone()
    idx[::3] = idx[::3].flip(1)  # hand the kernel some pairs in the wrong order: it must re-rank in fp32
    col_best = ((o["col_val"].view(torch.int32).long() ^ 0x80000000) << 32)  # placeholder order bits (positive sims)
    col_best = col_best | (0xFFFFFFFF - o["col_idx"])
    ridx = idx.to(torch.int32).cuda().contiguous()
    Xd, Yd, cbd = X.cuda(), Y.cuda(), col_best.cuda()  # keep the device copies alive across the launch
    d = torch.empty(n, 2, device="cuda")
    w = torch.empty(n, device="cuda")
    mu = torch.empty(n, dtype=torch.uint8, device="cuda")
    L.call("mv_k3_ratio_mutual", L.ptr(Xd), L.ptr(Yd), C, None, n, L.ptr(ridx), L.ptr(cbd),
           int(ratio_test), L.ptr(d), L.ptr(w), L.ptr(mu), stream())
    want_d = 1 - F.cosine_similarity(Y[o["row_idx"]], X[:, None, :], dim=-1)
    torch.testing.assert_close(d.cpu(), want_d, rtol=0, atol=1e-6)
    assert torch.equal(ridx.cpu().long(), o["row_idx"])
    want_w = restated.ratio_weights(want_d) if ratio_test else want_d[:, 0]
    torch.testing.assert_close(w.cpu(), want_w, rtol=0, atol=2e-4 if ratio_test else 1e-6)
    # the weight must be exactly the reference formula applied to the kernel's own distances
    if ratio_test:
        torch.testing.assert_close(w.cpu(), restated.ratio_weights(d.cpu()), rtol=0, atol=1e-7)
    assert torch.equal(mu.cpu().bool(), o["mutual"])


@pytest.mark.parametrize("n,k", [(1, 1), (5, 10), (1000, 1000), (19200, 1000), (12544, 500), (3000, 1), (70000, 16384), (1024, 1024)])
def test_topk_matches(mv, n, k):
    L = mv._lib
    g = torch.Generator().manual_seed(n + k)
    w = torch.randn(n, generator=g)
    w[n // 2:] = w[: n - n // 2].clone() if n > 3 else w[n // 2:]  # exact duplicates: ties
    if n > 10:
        w[3] = float("-inf")
        w[5] = 1e30
    idx = torch.randint(0, 1 << 20, (n, 2), generator=g, dtype=torch.int32)
    kk = min(k, n)
    src = torch.empty(kk, dtype=torch.int32, device="cuda")
    dst = torch.empty(kk, dtype=torch.int32, device="cuda")
    val = torch.empty(kk, device="cuda")
    kd = torch.zeros(1, dtype=torch.int32, device="cuda")
    wd, idxd = w.cuda(), idx.cuda()
    L.call("mv_k3_topk_matches", L.ptr(wd), L.ptr(idxd), None, n, k, L.ptr(src), L.ptr(dst), L.ptr(val), L.ptr(kd), stream())
    assert int(kd.item()) == kk
    want_v, _ = torch.topk(w, kk)
    assert torch.equal(val.cpu(), want_v)                       # same multiset, sorted descending
    s = src.cpu().long()
    assert torch.equal(w[s], val.cpu()) and s.unique().numel() == kk
    assert torch.equal(dst.cpu(), idx[s, 0])
    # ties resolved towards the lower row, and rows ascending inside a tie group
    same = val.cpu()[1:] == val.cpu()[:-1]
    assert (s[1:][same] > s[:-1][same]).all()
    if kk < n:
        thr = want_v[-1]
        tied_rows = (w == thr).nonzero().squeeze(1)
        taken = s[val.cpu() == thr]
        assert torch.equal(taken, tied_rows[: taken.numel()])


def test_topk_device_count(mv):
    L = mv._lib
    w = torch.arange(100, dtype=torch.float32)
    idx = torch.zeros(100, 2, dtype=torch.int32)
    nd = torch.tensor([40], dtype=torch.int32, device="cuda")
    src = torch.empty(10, dtype=torch.int32, device="cuda")
    dst = torch.empty(10, dtype=torch.int32, device="cuda")
    val = torch.empty(10, device="cuda")
    wd, idxd = w.cuda(), idx.cuda()
    L.call("mv_k3_topk_matches", L.ptr(wd), L.ptr(idxd), L.ptr(nd), 100, 10, L.ptr(src), L.ptr(dst), L.ptr(val), None, stream())
    assert src.cpu().tolist() == list(range(39, 29, -1))


@pytest.mark.parametrize("k", [1, 100, 1000, 5000])
def test_score_counts_equal_oracle_recall(mv, syn, k):
    ev = mv.evaluation
    g = torch.Generator().manual_seed(k)
    n0, n1 = 3000, 2500
    xyz0 = torch.randn(n0, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
    xyz1 = torch.randn(n1, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
    Rt = syn.random_rt(g, max_deg=10.0, t_sigma=0.01)
    Kmat = torch.tensor([[300.0, 0.0, 80.0], [0.0, 300.0, 60.0], [0.0, 0.0, 1.0]])
    src = torch.randint(0, n0, (k,), generator=g, dtype=torch.int32)
    dst = torch.randint(0, n1, (k,), generator=g, dtype=torch.int32)
    xyz1[dst.long()] = restated.transform_points_Rt(xyz0[src.long()], Rt) + 0.02 * torch.randn(k, 3, generator=g)
    mutual = (torch.rand(n0, generator=g) < 0.5).to(torch.uint8)
    thr3 = [0.01, 0.02, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5]
    thr2 = [1, 2, 5, 15, 25, 35, 50]
    acc = ev.RecallAccumulator(thr3, thr2, device="cuda")
    m = mv.correspondence.MatchResult()
    m.k, m.k_dev = k, None
    m.sel_src, m.sel_dst, m.mutual = src.cuda(), dst.cuda(), mutual.cuda()
    e3, e2 = acc.score(m, xyz0.cuda(), xyz1.cuda(), Rt, Kmat, want_errors=True)
    o3, o2 = restated.pair_errors(xyz0[src.long()], xyz1[dst.long()], Rt, Kmat)
    torch.testing.assert_close(e3.cpu()[:k], o3, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(e2.cpu()[:k], o2, rtol=1e-4, atol=1e-4)
    h = acc.hits.cpu().tolist()
    assert h[0] == k and h[1] == int(mutual[src.long()].sum())
    # integer counts == counts of the kernel's own errors (exact), and within 0.1 pp of the oracle's recall
    for i, t in enumerate(thr3):
        assert h[2 + i] == int((e3.cpu()[:k] < t).sum())
        assert abs(100.0 * h[2 + i] / k - restated.recall(o3, [t])[0]) <= 0.1 + 100.0 / k
    for i, t in enumerate(thr2):
        assert h[2 + len(thr3) + i] == int((e2.cpu()[:k] < t).sum())
    mu = mutual[src.long()].bool()
    assert h[2 + len(thr3) + len(thr2)] == int(((e3.cpu()[:k] < thr3[0]) & mu).sum())
    acc.score(m, xyz0.cuda(), xyz1.cuda(), Rt, Kmat)  # counters accumulate
    assert acc.hits.cpu().tolist() == [2 * v for v in h]


@pytest.mark.parametrize("shape", [(5, 7, 9), (30, 50, 50), (1, 1, 1), (20, 14, 14)])
def test_argmax_2d(mv, shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    x[0].flatten()[::2] = x[0].max() + 1  # ties: first occurrence
    for mx in (True, False):
        got = mv.correspondence.argmax_2d(x.cuda(), max_value=mx)
        assert got.dtype == torch.int64 and got.is_cuda
        assert torch.equal(got.cpu(), restated.argmax_2d(x, max_value=mx))
    assert mv.correspondence.argmax_2d(x).device.type == "cpu"


@pytest.mark.parametrize("idx", [1, 2, 3])
@pytest.mark.parametrize("shape", [dict(C=768, h=14, w=14, K=20, image_size=224), dict(C=64, h=50, w=50, K=30, image_size=800)])
def test_spair_errors(mv, syn, idx, shape):
    p = syn.spair_pair(idx, **shape)
    mv.correspondence.set_match_precision(dtype="tf32")
    try:
        hits = torch.zeros(2, dtype=torch.int64, device="cuda")
        es, en, isame, inn, pred = mv.spair.compute_errors_from_features(
            p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], p["image_size"], hits=hits, return_heatmap_argmax=True)
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    oes, oen, oisame, oinn, heat = restated.spair_compute_errors(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"],
                                                                 p["image_size"], return_pred=True)
    # arg-max identical wherever the oracle's top-2 heat-map gap exceeds 1e-3 (north-star tolerance)
    flat = heat.flatten(1)
    top2 = torch.topk(flat, 2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(pred[clear], flat.argmax(1)[clear])
    assert torch.equal(isame, oisame)
    ok = clear[oisame]
    torch.testing.assert_close(es[ok], oes[ok], rtol=0, atol=1e-5)
    torch.testing.assert_close(en[ok], oen[ok], rtol=0, atol=1e-5)
    assert torch.equal(inn[ok], oinn[ok])
    h = hits.cpu().tolist()
    assert h[0] == oisame.numel() and h[1] == int((es < 0.10).sum())


def test_spair_confusion_and_dataset_recall(mv, syn):
    """spair.evaluate_pairs == evaluate_dataset (evaluate_spair_correspondence.py:106-123) on the oracle's outputs."""
    pairs = [syn.spair_pair(i) for i in range(6)]
    mv.correspondence.set_match_precision(dtype="tf32")
    try:
        recall, conf = mv.spair.evaluate_pairs(pairs, pck_thresh=0.10)
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    errs, src, tgt = [], [], []
    for p in pairs:
        es, en, isame, inn = restated.spair_compute_errors(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], p["image_size"])
        errs.append(es); src.append(isame); tgt.append(inn)
    errs, src, tgt = torch.cat(errs), torch.cat(src), torch.cat(tgt)
    want = torch.zeros_like(conf)
    for a, b in zip(src.tolist(), tgt.tolist()):
        want[a, b] += 1
    assert int(conf.sum()) == src.numel()
    assert (conf - want).abs().sum() <= 2          # a near-tie in one heat map may move one entry
    assert abs(recall - 100.0 * (errs < 0.10).float().mean().item()) <= 100.0 / errs.numel() + 1e-6


def test_angle_binned_recall(mv, syn):
    """integer per-bin counts == compute_binned_performance on per-pair recall@2cm (equal k per pair)."""
    ev = mv.evaluation
    tr = mv.transformations
    g = torch.Generator().manual_seed(3)
    acc = ev.RecallAccumulator([0.01, 0.02, 0.05], [5, 25, 50], device="cuda", angle_bins=[0, 30, 60, 90, 120])
    k, n = 200, 1000
    rec, ang = [], []
    for i in range(12):
        Rt = syn.random_rt(g, max_deg=119.0, t_sigma=0.05)
        xyz0 = torch.randn(n, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
        src = torch.randperm(n, generator=g)[:k].to(torch.int32)
        dst = torch.randperm(n, generator=g)[:k].to(torch.int32)
        xyz1 = torch.randn(n, 3, generator=g)
        xyz1[dst.long()] = restated.transform_points_Rt(xyz0[src.long()], Rt) + 0.02 * torch.randn(k, 3, generator=g)
        m = mv.correspondence.MatchResult()
        m.k, m.k_dev, m.sel_src, m.sel_dst, m.mutual = k, None, src.cuda(), dst.cuda(), None
        e3, _ = acc.score(m, xyz0.cuda(), xyz1.cuda(), Rt, torch.eye(3), want_errors=True)
        rec.append((e3.cpu()[:k] < 0.02).float().mean())
        ang.append(tr.so3_rotation_angle(Rt[None, :3, :3])[0] * 180.0 / 3.141592653589793)
    want = mv.correspondence.compute_binned_performance(torch.stack(rec), torch.stack(ang), [0, 30, 60, 90, 120])
    got = acc.binned_recall()
    for a, b in zip(got, want):
        assert (a != a and b != b) or abs(a - 100.0 * float(b)) < 1e-3
    torch.testing.assert_close(tr.so3_rotation_angle(Rt[None, :3, :3]), tr.so3_relative_angle(Rt[None, :3, :3], torch.eye(3)[None]))


@pytest.mark.parametrize("shape,B", [(dict(C=768, h=14, w=14, K=20, image_size=224), 7), (dict(C=64, h=50, w=50, K=30, image_size=800), 3),
                                     (dict(C=40, h=5, w=7, K=3, image_size=64), 4), (dict(C=3072, h=14, w=14, K=9, image_size=224), 2),
                                     # streaming kernel: 256 pixels (all eight consumer warps), an odd number of chunks with a
                                     # 24-channel tail, two key-point tiles (K > 32); an odd pixel count with 4-byte A loads
                                     (dict(C=72, h=16, w=16, K=40, image_size=256), 3), (dict(C=104, h=15, w=15, K=17, image_size=240), 3)])
def test_spair_batch_equals_oracle_per_pair(mv, syn, shape, B):
    """mv_spair_match_batch (one launch for B pairs, fp32) against the oracle pair by pair: arg-max identical
    wherever the oracle's top-2 heat-map gap exceeds 1e-5 (fp32 summation order only), errors to 1e-5, the
    key points kept (in both images) bit-exact, integer PCK counts and confusion equal to the oracle's."""
    pairs = [syn.spair_pair(20 + i, **shape) for i in range(B)]
    K = shape["K"]
    hits = torch.zeros(2, dtype=torch.int64, device="cuda")
    conf = torch.zeros((K, K), dtype=torch.int64, device="cuda")
    es, en, inn, pred = mv.spair.compute_errors_batch(
        torch.stack([p["feats"] for p in pairs]), torch.stack([p["kps_i"] for p in pairs]), torch.stack([p["kps_j"] for p in pairs]),
        [p["thresh_scale"] for p in pairs], shape["image_size"], hits=hits, confusion=conf)
    es, en, inn, pred = es.cpu(), en.cpu(), inn.cpu().long(), pred.cpu().long()
    n_both = n_hit = 0
    want_conf = torch.zeros((K, K), dtype=torch.int64)
    unclear = 0
    for b, p in enumerate(pairs):
        oes, oen, oisame, oinn, heat = restated.spair_compute_errors(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"],
                                                                     p["image_size"], return_pred=True)
        flat = heat.flatten(1)
        top2 = torch.topk(flat, 2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-5
        unclear += int((~clear).sum())
        assert torch.equal(pred[b][clear], flat.argmax(1)[clear])
        kept = (es[b] >= 0).nonzero().squeeze(1)
        assert torch.equal(kept, oisame)                                  # which key points are in both: bit-exact
        assert torch.equal((en[b] >= 0), (es[b] >= 0)) and torch.equal((inn[b] >= 0), (es[b] >= 0))
        ok = clear[oisame]
        torch.testing.assert_close(es[b][oisame][ok], oes[ok], rtol=0, atol=1e-5)
        if bool(clear.all()):
            torch.testing.assert_close(en[b][oisame], oen, rtol=0, atol=1e-5)
            assert torch.equal(inn[b][oisame], oinn)
            for a_, b_ in zip(oisame.tolist(), oinn.tolist()):
                want_conf[a_, b_] += 1
        else:
            for a_, b_ in zip(oisame.tolist(), inn[b][oisame].tolist()):
                want_conf[a_, b_] += 1
        n_both += oisame.numel()
        n_hit += int((es[b][oisame] < 0.10).sum())
    assert unclear <= B  # the tolerance set is almost everything
    assert hits.cpu().tolist() == [n_both, n_hit]
    assert torch.equal(conf.cpu(), want_conf)


def test_spair_batch_equals_per_pair_path_and_edge_cases(mv, syn):
    """the batched launch and the per-pair path (kernels 1-3, tf32) agree; B = 0, K = 0 and a pair without any
    common key point are handled; evaluate_batches == evaluate_pairs."""
    sp = mv.spair
    pairs = [syn.spair_pair(40 + i) for i in range(5)]
    pairs[2]["kps_j"][:, 2] = 0.0  # nothing in both images
    feats = torch.stack([p["feats"] for p in pairs])
    ki, kj = torch.stack([p["kps_i"] for p in pairs]), torch.stack([p["kps_j"] for p in pairs])
    ts = [p["thresh_scale"] for p in pairs]
    es, en, inn, pred = sp.compute_errors_batch(feats, ki, kj, ts, 224)
    assert bool((es[2] == -1).all()) and bool((inn[2] == -1).all())
    mv.correspondence.set_match_precision(dtype="tf32")
    try:
        for b, p in enumerate(pairs):
            a = sp.compute_errors_from_features(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], 224, return_heatmap_argmax=True)
            same = a[4].cpu() == pred[b].cpu().long()
            assert same.float().mean() >= 0.9            # tf32 ranking vs fp32: only near-ties may differ
            keep = (es[b] >= 0).nonzero().squeeze(1).cpu()
            assert torch.equal(a[2], keep)
        r1, c1 = sp.evaluate_pairs(pairs, kp_max=30)
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    r2, c2 = sp.evaluate_batches([dict(feats=feats, kps_i=ki, kps_j=kj, thresh_scale=ts, image_size=224)], kp_max=30)
    assert int(c1.sum()) == int(c2.sum()) and (c1 - c2).abs().sum() <= 4 and abs(r1 - r2) <= 100.0 * 2 / max(int(c1.sum()), 1) + 1e-6
    # degenerate sizes
    e = sp.compute_errors_batch(feats[:0], ki[:0], kj[:0], [], 224)
    assert e[0].shape == (0, 20)
    e = sp.compute_errors_batch(feats, ki[:, :0], kj[:, :0], ts, 224)
    assert e[0].shape == (5, 0)
    with pytest.raises(ValueError):
        sp.compute_errors_batch(feats[0], ki, kj, ts, 224)


def test_spair_batch_many_pairs_per_cta(mv, syn):
    """more pairs than resident CTAs (2 per SM): every CTA of the streaming kernel walks several pairs, its producer warp
    running ahead across pair boundaries (ring phases, q buffers and the score matrix carry over).  Pair b of the long
    batch must equal, bit for bit, the same pair in a batch small enough for one pair per CTA; the short batch is
    checked against the oracle."""
    shape = dict(C=48, h=14, w=14, K=20, image_size=224)
    distinct = 37
    pairs = [syn.spair_pair(100 + i, **shape) for i in range(distinct)]
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = 4 * sm + 45  # ragged: some CTAs take one pair more than others
    pick = lambda key: torch.stack([pairs[b % distinct][key] for b in range(B)])
    ts = [pairs[b % distinct]["thresh_scale"] for b in range(B)]
    hits = torch.zeros(2, dtype=torch.int64, device="cuda")
    long = mv.spair.compute_errors_batch(pick("feats"), pick("kps_i"), pick("kps_j"), ts, 224, hits=hits)
    hits_s = torch.zeros(2, dtype=torch.int64, device="cuda")
    short = mv.spair.compute_errors_batch(torch.stack([p["feats"] for p in pairs]), torch.stack([p["kps_i"] for p in pairs]),
                                          torch.stack([p["kps_j"] for p in pairs]), ts[:distinct], 224, hits=hits_s)
    idx = torch.arange(B, device="cuda") % distinct
    for a, b in zip(long, short):
        assert torch.equal(a, b[idx])
    reps = torch.bincount(idx.cpu(), minlength=distinct)
    per_pair_both = (short[0] >= 0).sum(1).cpu()
    assert int(hits[0]) == int((per_pair_both * reps).sum())
    for b, p in enumerate(pairs):
        _, _, oisame, _, heat = restated.spair_compute_errors(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], 224, return_pred=True)
        flat = heat.flatten(1)
        top2 = torch.topk(flat, 2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-5
        assert torch.equal(short[3][b].cpu().long()[clear], flat.argmax(1)[clear])
        assert torch.equal((short[0][b].cpu() >= 0).nonzero().squeeze(1), oisame)


def test_spair_batch_single_tf32_term(mv, syn):
    """spair.set_heatmap_precision("tf32"): one tf32 MMA per tile instead of three.  Arg-max identical to the oracle's
    wherever its top-2 heat-map gap exceeds 1e-3 (the tf32 tolerance of the matching path), key points kept bit-exact,
    PCK within one key point per pair of the default mode; the default mode is back afterwards."""
    shape = dict(C=768, h=14, w=14, K=20, image_size=224)
    pairs = [syn.spair_pair(60 + i, **shape) for i in range(12)]
    args = (torch.stack([p["feats"] for p in pairs]), torch.stack([p["kps_i"] for p in pairs]), torch.stack([p["kps_j"] for p in pairs]),
            [p["thresh_scale"] for p in pairs], 224)
    h3 = torch.zeros(2, dtype=torch.int64, device="cuda")
    ref = mv.spair.compute_errors_batch(*args, hits=h3)
    assert mv.spair.set_heatmap_precision("tf32") == "3xtf32"
    try:
        h1 = torch.zeros(2, dtype=torch.int64, device="cuda")
        es, en, inn, pred = mv.spair.compute_errors_batch(*args, hits=h1)
    finally:
        assert mv.spair.set_heatmap_precision("3xtf32") == "tf32"
    compared = total = 0
    for b, p in enumerate(pairs):
        _, _, oisame, _, heat = restated.spair_compute_errors(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], 224, return_pred=True)
        flat = heat.flatten(1)
        top2 = torch.topk(flat, 2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-3
        compared += int(clear.sum())
        total += clear.numel()
        assert torch.equal(pred[b].cpu().long()[clear], flat.argmax(1)[clear])
        assert torch.equal((es[b].cpu() >= 0).nonzero().squeeze(1), oisame)
    assert compared >= 0.5 * total, (compared, total)
    assert int(h1[0]) == int(h3[0]) and abs(int(h1[1]) - int(h3[1])) <= len(pairs)
    again = mv.spair.compute_errors_batch(*args)
    assert torch.equal(again[3], ref[3])


def test_spair_paths_vs_reference_golden(mv, syn, golden):
    """both SPair paths against tests/golden/spair_small.npz, the output of the reference's own compute_errors."""
    from oracle.make_golden import SPAIR_SMALL

    g = golden("spair_small")
    p = syn.spair_pair(int(g["index"]), **SPAIR_SMALL)
    want_same, want_nn = torch.from_numpy(g["error_same"]), torch.from_numpy(g["error_nn"])
    want_isame, want_inn = torch.from_numpy(g["index_same"]).long(), torch.from_numpy(g["index_nn"]).long()
    es, en, inn, pred = mv.spair.compute_errors_batch(p["feats"][None], p["kps_i"][None], p["kps_j"][None], [p["thresh_scale"]],
                                                      p["image_size"])
    es, en, inn = es[0].cpu(), en[0].cpu(), inn[0].cpu().long()
    keep = (es >= 0).nonzero().squeeze(1)
    assert torch.equal(keep, want_isame)
    torch.testing.assert_close(es[keep], want_same, rtol=0, atol=1e-5)
    torch.testing.assert_close(en[keep], want_nn, rtol=0, atol=1e-5)
    assert torch.equal(inn[keep], want_inn)
    w = p["feats"].shape[-1]
    pr = pred[0].cpu().long()
    assert torch.equal(torch.stack((pr % w, pr // w), dim=1), torch.from_numpy(g["pred"]).long())
    mv.correspondence.set_match_precision(dtype="tf32")
    try:
        a = mv.spair.compute_errors_from_features(p["feats"], p["kps_i"], p["kps_j"], p["thresh_scale"], p["image_size"])
    finally:
        mv.correspondence.set_match_precision(dtype=DEFAULT_DTYPE)
    assert torch.equal(a[2], want_isame)
    same = a[3] == want_inn
    assert same.float().mean() >= 0.9   # tf32 ranking: only near-ties of the heat map may differ
