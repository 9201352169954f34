"""Kernel 1 and its producers against the oracle: index lists and tap origins bit-exact, values to 1e-5."""
from ctypes import c_float, c_void_p

import pytest
import torch
import torch.nn.functional as F

from oracle import restated

pytestmark = pytest.mark.gpu
ATOL = 1e-5  # fp32 interpolation + normalisation, stated tolerance for kernel-1 values


@pytest.fixture(scope="module")
def C_(mv):
    return mv.correspondence


@pytest.mark.parametrize("C,h,w", [(8, 1, 1), (64, 6, 8), (768, 14, 14), (2048, 15, 20), (3072, 28, 28), (40, 5, 3)])
def test_chw_to_hwc_and_pixel_norm(C_, C, h, w):
    g = torch.Generator().manual_seed(C + h)
    f = torch.randn(C, h, w, generator=g)
    f[:, 0, 0] = 0.0  # a zero pixel exercises the 1e-12 clamp
    d = f.cuda()
    out = C_._hwc(d).cpu()
    assert torch.equal(out, f.reshape(C, -1).t().contiguous())
    outn = C_._hwc(d, prenorm=True).cpu()
    ref = F.normalize(f, p=2, dim=0).reshape(C, -1).t()
    torch.testing.assert_close(outn, ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,stride,p_valid", [(1, 1, 1.0), (1, 1, 0.0), (777, 3, 0.5), (19200, 3, 0.95), (12544, 1, 0.4),
                                               (4096, 1, 0.0), (4096, 1, 1.0), (100000, 1, 0.3)])
def test_compact_valid_is_bit_exact(mv, n, stride, p_valid):
    L = mv._lib
    g = torch.Generator().manual_seed(n)
    z = torch.rand(n, stride, generator=g)
    z[torch.rand(n, generator=g) >= p_valid, stride - 1] = 0.0
    z[::7, stride - 1] *= -1.0  # negatives are not valid either
    d = z.cuda()
    idx = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("mv_compact_valid", c_void_p(d.data_ptr() + 4 * (stride - 1)), stride, n, L.ptr(idx), L.ptr(cnt),
           c_void_p(torch.cuda.current_stream().cuda_stream))
    want = (z[:, stride - 1] > 0).nonzero().squeeze(1).to(torch.int32)
    assert int(cnt.item()) == want.numel()
    assert torch.equal(idx[: want.numel()].cpu(), want)


@pytest.mark.parametrize("shape", [dict(C=64, h=6, w=8, H=24, W=32), dict(C=2048, h=15, w=20, H=120, W=160),
                                   dict(C=768, h=6, w=8, H=48, W=64), dict(C=1024, h=6, w=8, H=48, W=64),
                                   dict(C=1536, h=6, w=8, H=24, W=32), dict(C=4096, h=6, w=8, H=48, W=64)])
def test_depth_side_matches_oracle(C_, mv, syn, shape):
    p = syn.scannet_pair(3, coherent=False, **shape)
    K = p["K"]
    xyz_o, f_o, keep_o = restated.depth_side(p["feat_0"], p["depth_0"], K)
    Kh, Kinv = C_._host_mat(K), C_._host_mat(K.inverse())
    s = C_.prepare_depth_side(p["feat_0"], p["depth_0"], Kh, Kinv, torch.device("cuda"), want_taps=True, rows="f32")
    n = s.n
    assert n == xyz_o.shape[0]
    assert torch.equal(s.valid_idx[:n].cpu().long(), keep_o)                       # gather indices: bit-exact
    torch.testing.assert_close(s.xyz[:n].cpu(), xyz_o, rtol=0, atol=1e-6)
    coords = restated.depth_side_coords(K, xyz_o, p["depth_0"].shape[-2:], p["feat_0"].shape[-2:])
    assert torch.equal(s.taps[:n].cpu().long(), torch.floor(coords).long())        # upsample tap origins: bit-exact
    ref = F.normalize(f_o, dim=-1)
    torch.testing.assert_close(s.rows32[:n].cpu(), ref, rtol=0, atol=ATOL)
    if s.rows16 is not None:  # the 16-bit operand rows (fp16 + 8 augmentation columns by default, query role: not centred)
        got16 = s.rows16[:n, :shape["C"]].float().cpu()
        assert (got16 - s.rows32[:n].cpu()).abs().max() <= 2 ** -8 * s.rows32[:n].abs().max().item() + 1e-8


@pytest.mark.parametrize("shape", [dict(C=64, h=8, w=8, H=32, W=32, radius=12.0), dict(C=3072, h=28, w=28, H=112, W=112, radius=40.0),
                                   dict(C=768, h=7, w=9, H=30, W=41, radius=14.0), dict(C=1024, h=9, w=9, H=36, W=36, radius=14.0),
                                   dict(C=1536, h=8, w=8, H=32, W=32, radius=12.0), dict(C=4096, h=8, w=8, H=16, W=16, radius=7.0),
                                   dict(C=2048, h=8, w=8, H=64, W=64, radius=30.0)])
def test_xyz_side_matches_oracle(C_, syn, shape):
    p = syn.navi_pair(5, coherent=False, **shape)
    xyz_o, f_o, uv_o, keep_o = restated.xyz_side(p["feat_0"], p["xyz_grid_0"])
    s = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], torch.device("cuda"), want_taps=True, rows="f32")
    n = s.n
    assert n == xyz_o.shape[0]
    assert torch.equal(s.valid_idx[:n].cpu().long(), keep_o)
    assert torch.equal(s.xyz[:n].cpu(), xyz_o)
    assert torch.equal(s.uv[:n].cpu(), uv_o)
    H, W = p["xyz_grid_0"].shape[-2:]
    h, w = p["feat_0"].shape[-2:]
    ys, xs = keep_o // W, keep_o % W
    fx = torch.floor((w / W) * (xs.float() + 0.5) - 0.5).long()
    fy = torch.floor((h / H) * (ys.float() + 0.5) - 0.5).long()
    assert torch.equal(s.taps[:n].cpu().long(), torch.stack((fx, fy), dim=1))
    torch.testing.assert_close(s.rows32[:n].cpu(), F.normalize(f_o, dim=-1), rtol=0, atol=ATOL)


def test_no_sync_variant_equals_synced(C_, syn):
    p = syn.navi_pair(6, coherent=False, C=64, h=8, w=8, H=32, W=32, radius=12.0)
    a = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], torch.device("cuda"), sync=True, rows="f32")
    b = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], torch.device("cuda"), sync=False, rows="f32")
    n = a.n
    assert int(b.n_dev.item()) == n and b.n == 32 * 32
    assert torch.equal(a.rows32[:n], b.rows32[:n]) and torch.equal(a.xyz[:n], b.xyz[:n])


@pytest.mark.parametrize("C,h,w,K", [(768, 14, 14, 20), (64, 50, 50, 30), (8, 2, 3, 1)])
def test_keypoint_sampling_align_corners_true(mv, C_, C, h, w, K):
    L = mv._lib
    g = torch.Generator().manual_seed(K)
    feats = torch.randn(C, h, w, generator=g)
    size = 224.0
    kps = torch.zeros(K, 3)
    kps[:, :2] = torch.randint(0, 224, (K, 2), generator=g).float()
    kps[0, :2] = torch.tensor([0.0, 223.0])
    fn = F.normalize(feats, p=2, dim=0)
    ndc = (kps[:, :2] / size * 2 - 1)[None, None]
    want = F.grid_sample(fn[None], ndc, mode="bilinear", align_corners=True)[0, :, 0].t()
    src = C_._hwc(feats.cuda(), prenorm=True)
    coords = torch.empty(K, 2, device="cuda")
    kd = kps.cuda()
    L.call("mv_geom_keypoint_coords", L.ptr(kd), 3, K, c_float(size), h, w, L.ptr(coords), C_._stream())
    _, got, _ = C_._sample(L.MV_SAMPLE_BILINEAR_ZEROS, src, C, h, w, coords, None, K, False, False, True)
    torch.testing.assert_close(got[:K].cpu(), want, rtol=0, atol=ATOL)


def test_sample_pointcloud_features_and_grid_to_pointcloud(C_, syn):
    p = syn.scannet_pair(9, coherent=False, C=64, h=6, w=8, H=24, W=32)
    K = p["K"]
    pc_o = restated.backproject(K.inverse(), p["depth_0"])
    pc = C_.grid_to_pointcloud(K.inverse(), p["depth_0"])
    assert pc.device.type == "cpu"
    torch.testing.assert_close(pc, pc_o, rtol=0, atol=1e-6)
    keep = pc_o[:, 2] > 0
    f = C_.sample_pointcloud_features(p["feat_0"], K.clone(), pc_o[keep], p["depth_0"].shape[-2:])
    want = restated.sample_pointcloud_features(p["feat_0"], K.clone(), pc_o[keep], p["depth_0"].shape[-2:])
    torch.testing.assert_close(f, want, rtol=0, atol=ATOL)


@pytest.mark.parametrize("C", [768, 2048, 3072, 200])
def test_split_rows_rebuild_the_fp32_rows(C_, syn, C):
    """kernel 1's split output (bf16 hi + bf16 residual) against its own fp32 rows: hi is the bf16 rounding of
    the fp32 row bit for bit, and hi + lo rebuilds it to 2^-16 relative (the stated tolerance of the format)."""
    p = syn.navi_pair(7, coherent=False, C=C, h=10, w=10, H=40, W=40, radius=15.0)
    dev = torch.device("cuda")
    C_.set_match_precision(dtype="bf16")
    try:
        a = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], dev, rows="f32")
        b = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], dev, rows="split")
    finally:
        C_.set_match_precision(dtype=C_.DEFAULT_DTYPE)
    n = a.n
    assert b.rows32 is None and b.rows_lo is not None
    assert torch.equal(a.rows16[:n], b.rows16[:n])
    assert torch.equal(a.rows32[:n].to(torch.bfloat16), b.rows16[:n])
    rebuilt = b.rows16[:n].float() + b.rows_lo[:n].float()
    err = (rebuilt - a.rows32[:n]).abs()
    assert bool((err <= 2.0 ** -16 * a.rows32[:n].abs() + 1e-30).all())
    # the default format: fp16 hi (C + 8 columns, 128-byte pitch) + fp16 residual scaled by 2^11 -> 2^-21 relative
    a = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], dev, rows="f32")
    b = C_.prepare_xyz_side(p["feat_0"], p["xyz_grid_0"], dev, rows="split")
    assert b.rows16.dtype == torch.float16 and b.rows16.shape[1] == (C + 8 + 63) // 64 * 64 and b.rows32 is None
    assert torch.equal(a.rows16[:n, :C + 8], b.rows16[:n, :C + 8])
    assert torch.equal(a.rows32[:n].to(torch.float16), b.rows16[:n, :C])
    rebuilt = b.rows16[:n, :C].float() + b.rows_lo[:n].float() / 2048.0
    err = (rebuilt - a.rows32[:n]).abs()
    assert bool((err <= 2.0 ** -21 * a.rows32[:n].abs() + 1e-9).all())


@pytest.mark.parametrize("shape", [dict(C=3072, h=28, w=28, H=112, W=112, radius=40.0), dict(C=768, h=14, w=14, H=56, W=56, radius=22.0),
                                   dict(C=1536, h=8, w=8, H=32, W=32, radius=13.0), dict(C=1024, h=6, w=6, H=48, W=48, radius=20.0),
                                   dict(C=2048, h=8, w=8, H=32, W=32, radius=15.0), dict(C=4096, h=4, w=4, H=16, W=16, radius=7.0),
                                   dict(C=768, h=8, w=8, H=16, W=32, radius=9.0), dict(C=384, h=5, w=7, H=5, W=28, radius=20.0)])
@pytest.mark.parametrize("role", ["query", "target"])
def test_grid_kernel_equals_point_run_kernel_and_oracle(C_, mv, syn, shape, role):
    """the tiled cluster kernel (csrc/k1_grid.cu: 2-D tiles, channels split over a thread-block cluster, partial sums of
    squares exchanged through distributed shared memory) against the point-run kernel on the same inputs and against the
    oracle: same arithmetic per element, only the order of the sum of squares differs."""
    L = mv._lib
    C, H, W = shape["C"], shape["H"], shape["W"]
    assert L.load().mv_k1_grid_supported(C, shape["h"], shape["w"], H, W) == 1
    p = syn.navi_pair(9, coherent=False, **shape)
    dev = torch.device("cuda")
    fm = C_._feature_map(p["feat_0"], dev)
    mu = C_._center(fm[0], fm[0].shape[0])
    kw = {"role": L.MV_ROLE_QUERY, "dotvec": mu} if role == "query" else {"role": L.MV_ROLE_TARGET, "center": mu}
    outs = {}
    for grid in (0, 1):
        C_.set_match_precision(k1_grid=grid)
        try:
            s = C_.prepare_xyz_side(fm, p["xyz_grid_0"], dev, **kw)
        finally:
            C_.set_match_precision(k1_grid=0)
        torch.cuda.synchronize()
        outs[grid] = s
    a, b = outs[0], outs[1]
    n = a.n
    assert b.n == n and n > 50 and b.rows16.shape == a.rows16.shape
    ha, hb = a.rows16[:n, :C].float(), b.rows16[:n, :C].float()
    ulp = (ha.abs().clamp(min=2.0 ** -14) * 2.0 ** -10)
    assert ((ha - hb).abs() <= ulp).all()                       # at most one fp16 step, on the few elements that sit at a rounding boundary
    assert (ha != hb).float().mean() < 0.01
    ra, rb = ha + a.rows_lo[:n].float() / 2048.0, hb + b.rows_lo[:n].float() / 2048.0
    full = ra + (mu[None] if role == "target" else 0.0)
    # the rebuilt fp32 rows: two hi / lo splits of values that differ by an ulp of the norm (2 x 2^-22 of the element + that ulp)
    assert ((ra - rb).abs() <= 1e-6 * full.abs() + 2e-9).all()
    auga, augb = a.rows16[:n, C:C + 8].float(), b.rows16[:n, C:C + 8].float()
    if role == "target":
        assert torch.equal(auga, augb)
    else:
        ra_, rb_ = auga[:, 0] + auga[:, 1] + auga[:, 2] / 2048.0, augb[:, 0] + augb[:, 1] + augb[:, 2] / 2048.0
        torch.testing.assert_close(ra_, rb_, rtol=0, atol=2e-6)
    _, f_o, _, _ = restated.xyz_side(p["feat_0"], p["xyz_grid_0"])
    want = F.normalize(f_o, dim=-1)
    rebuilt = rb + (mu[None] if role == "target" else 0.0)
    torch.testing.assert_close(rebuilt.cpu(), want, rtol=0, atol=ATOL)
