"""The f16c operand form (include/mvmatch.h, mv_k1_sample_f16c): fp16 rows with the target set stored relative to
a centre and the queries carrying their dot product with that centre in three augmentation columns.

Checked here: the row planes bit for bit against torch's own fp16 rounding, the centre, that kernel 2 on such
rows returns a . b (not a . (b - mu)) for rows AND columns, and that on nearly collinear all-positive features
(the regime of CNN feature maps, where a plain bf16 product ranks the wrong neighbour on most rows) the
neighbours equal the fp64 ranking wherever its top-2 gap exceeds 1e-5 -- 100x tighter than the north-star rule."""
from ctypes import c_size_t

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def collinear_rows(n, C, seed, spread=0.25):
    """unit rows = a common positive direction + a small individual part (mean cosine ~ 1 - spread^2)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(C, generator=g) + 0.2
    x = base[None, :] / base.norm() + spread * torch.randn(n, C, generator=g) / C ** 0.5
    return F.normalize(x.abs(), dim=1)


def f16c_rows(mv, X, role, center=None, dotvec=None, normalize=True, lo=True):
    L, C_ = mv._lib, mv.correspondence
    Xd = X.cuda().contiguous()
    n, C = Xd.shape
    hi = torch.full((n, L.f16c_pitch(C)), 9.0, dtype=torch.float16, device="cuda")
    lo_t = torch.empty((n, C), dtype=torch.float16, device="cuda") if lo else None
    rdot = torch.empty(n, device="cuda")
    L.call("mv_k1_sample_f16c", L.MV_SAMPLE_ROWS, L.ptr(Xd), C, 0, 0, None, None, n, int(normalize), role, L.ptr(center),
           L.ptr(dotvec), None, L.ptr(hi), hi.shape[1], L.ptr(lo_t), None, L.ptr(rdot), None, C_._stream())
    torch.cuda.synchronize()
    return hi, lo_t, rdot


@pytest.mark.parametrize("C", [768, 2048, 3072, 200])  # 200: the generic (any C) instantiation
def test_f16c_row_planes_bit_exact(mv, C):
    L, C_ = mv._lib, mv.correspondence
    X = collinear_rows(300, C, 1) * 3.0  # not unit: kernel 1 normalises
    mu = C_._center(X.cuda().contiguous(), 300)
    Xn = F.normalize(X.cuda(), dim=1)
    want_mu = Xn.mean(0)
    torch.testing.assert_close(mu, want_mu, rtol=1e-5, atol=1e-7)
    # target role: hi = fp16(x - mu), lo = fp16((x - mu - hi) * 2048), aug = (1, 1, 2^-11, 0...)
    hi, lo, _ = f16c_rows(mv, X, L.MV_ROLE_TARGET, center=mu)
    # kernel 1's own fp32 normalisation: x * (1 / max(||x||, eps)) -- rebuild it the same way
    ss = (X.cuda() * X.cuda()).sum(1, keepdim=True)
    y = X.cuda() * (1.0 / ss.sqrt().clamp(min=1e-12)) - mu[None]
    got = hi[:, :C].float()
    # the sum of squares is accumulated in a different order than torch's: allow one fp16 ulp on a handful of elements
    ulp = (y.abs().clamp(min=2.0 ** -14) * 2.0 ** -10)
    assert ((got - y).abs() <= 0.5 * ulp * 1.01 + 2e-8).all()  # + a few fp32 ulps of the normalised row (summation order)
    assert (got != y.half().float()).float().mean() < 2e-3
    rebuilt = got + lo.float() / 2048.0
    assert ((rebuilt - y).abs() <= y.abs() * 2.0 ** -21 + 2e-8).all()
    aug = hi[:, C:C + 8].float().cpu()
    assert torch.equal(aug, torch.tensor([1.0, 1.0, 2.0 ** -11, 0, 0, 0, 0, 0]).expand(300, 8))
    assert (hi[:, C + 8:] == 9.0).all()  # the padding up to the 128-byte pitch is never written (and never read by kernel 2)
    # query role: plain fp16 rows, aug = three pieces of r = x . mu
    hq, lq, r = f16c_rows(mv, X, L.MV_ROLE_QUERY, dotvec=mu)
    want_r = (Xn * mu[None]).sum(1)
    torch.testing.assert_close(r, want_r, rtol=0, atol=2e-6)
    a = hq[:, C:C + 8].float()
    torch.testing.assert_close(a[:, 0] + a[:, 1] + a[:, 2] / 2048.0, r, rtol=0, atol=1e-7)
    assert (a[:, 3:] == 0).all()
    assert ((hq[:, :C].float() - Xn).abs() <= Xn.abs() * 2.0 ** -11 * 1.01 + 1e-7).all()


@pytest.mark.parametrize("C,cluster", [(768, -1), (2048, -1), (2048, 0), (3072, 20), (264, 2)])
def test_f16c_product_is_the_uncentred_product(mv, C, cluster):
    """kernel 2 on f16c rows: row values / indices and column arg-max are those of a . b."""
    L, C_ = mv._lib, mv.correspondence
    n, m = 700, 900
    X, Y = collinear_rows(n, C, 2), collinear_rows(m, C, 3)
    mu = C_._center(Y.cuda().contiguous(), m)
    A, _, _ = f16c_rows(mv, X, L.MV_ROLE_QUERY, dotvec=mu, lo=False)
    B, _, _ = f16c_rows(mv, Y, L.MV_ROLE_TARGET, center=mu, lo=False)
    row_val = torch.empty((n, 2), device="cuda")
    row_idx = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    col_best = torch.empty(m, dtype=torch.int64, device="cuda")
    wsb = L.load().mv_k2_workspace_bytes(n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    L.call("mv_k2_sim_top2_ld", L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], n, m, C + 8, None, None, L.MV_DTYPE_F16, cluster, L.ptr(row_val), L.ptr(row_idx),
           L.ptr(col_best), L.ptr(ws), c_size_t(wsb), C_._stream())
    col_val = torch.empty(m, device="cuda")
    col_idx = torch.empty(m, dtype=torch.int32, device="cuda")
    L.call("mv_k2_unpack_col", L.ptr(col_best), m, L.ptr(col_val), L.ptr(col_idx), C_._stream())
    torch.cuda.synchronize()
    S = F.normalize(X.double(), dim=1) @ F.normalize(Y.double(), dim=1).t()
    val, idx = torch.topk(S, 3, dim=1)
    assert (row_val.cpu().double() - val[:, :2]).abs().max() < 4e-5      # a plain bf16 product is off by 2-7e-4 here
    clear = (val[:, 0] - val[:, 1]) > 5e-5
    assert clear.float().mean() > 0.5
    assert torch.equal(row_idx.cpu().long()[clear, 0], idx[clear, 0])
    clear2 = clear & ((val[:, 1] - val[:, 2]) > 5e-5)
    assert torch.equal(row_idx.cpu().long()[clear2, 1], idx[clear2, 1])
    cs = torch.topk(S, 2, dim=0)
    cclear = (cs.values[0] - cs.values[1]) > 5e-5
    assert torch.equal(col_idx.cpu().long()[cclear], cs.indices[0][cclear])
    assert (col_val.cpu().double() - cs.values[0]).abs().max() < 4e-5


def test_f16c_beats_plain_16bit_products_on_collinear_rows(mv):
    """the reason the form exists: on all-positive, nearly collinear rows a bf16 product ranks the wrong neighbour on
    most rows, the f16c product agrees with fp64 wherever its top-2 gap exceeds 1e-5."""
    C_ = mv.correspondence
    X, Y = collinear_rows(2000, 2048, 4, 0.1), collinear_rows(2500, 2048, 5, 0.1)
    S = F.normalize(X.double(), dim=1) @ F.normalize(Y.double(), dim=1).t()
    val, idx = torch.topk(S, 2, dim=1)
    clear = (val[:, 0] - val[:, 1]) > 1e-5
    agree = {}
    for dt in ("f16", "bf16"):
        C_.set_match_precision(dtype=dt)
        try:
            _, i = C_.knn_points(X, Y, 2, "cosine")
        finally:
            C_.set_match_precision(dtype=C_.DEFAULT_DTYPE)
        agree[dt] = float((i[clear, 0] == idx[clear, 0]).float().mean())
    print("top-1 agreement with fp64 on rows with gap > 1e-5:", agree)
    assert agree["f16"] == 1.0
    # documents the failure mode the default avoids: bf16 proposes a wrong pair of candidates on some of these rows even
    # after kernel 3's fp32 re-rank of the two (measured 0.983; 0.87 for its raw arg-max in a CPU simulation of the roundings)
    assert agree["bf16"] < 0.995


@pytest.mark.parametrize("kind", ["navi", "scannet"])
def test_pixel_dot_form_of_the_query_rows(mv, syn, kind):
    """the query rows' r = row . mu from the per-source-pixel dots (the row's own 4- / 16-tap blend of src[p] . mu, times
    1 / norm: nothing in kernel 1's hot loop) against the C-long dot product per row: the same number up to fp32 summation
    order, everything else bit-identical."""
    C_, L = mv.correspondence, mv._lib
    dev = torch.device("cuda")
    if kind == "navi":
        p = syn.navi_pair(3, coherent=False, C=768, h=14, w=14, H=56, W=56, radius=22.0)
        prep = lambda fm, kw: C_.prepare_xyz_side(fm, p["xyz_grid_0"], dev, **kw)
    else:
        p = syn.scannet_pair(3, coherent=False, C=2048, h=15, w=20, H=60, W=80)
        Kh, Kinv = C_._host_mat(p["K"]), C_._host_mat(p["K"].inverse())
        prep = lambda fm, kw: C_.prepare_depth_side(fm, p["depth_0"], Kh, Kinv, dev, **kw)
    fm0, fm1, kw0, _ = C_._pair_maps(p["feat_0"], p["feat_1"], dev)
    assert kw0["pixdot"] is not None
    a = prep(fm0, kw0)                                                  # pixel-dot form (the default)
    b = prep(fm0, {"role": kw0["role"], "dotvec": kw0["dotvec"]})       # dot product inside the row loop
    n, C = a.n, fm0[1]
    assert torch.equal(a.rows16[:n, :C], b.rows16[:n, :C]) and torch.equal(a.rows_lo[:n], b.rows_lo[:n])
    ra = a.rows16[:n, C:C + 3].float()
    rb = b.rows16[:n, C:C + 3].float()
    ra, rb = ra[:, 0] + ra[:, 1] + ra[:, 2] / 2048.0, rb[:, 0] + rb[:, 1] + rb[:, 2] / 2048.0
    assert (ra - rb).abs().max() <= 2e-6 and ra.abs().max() > 1e-5
