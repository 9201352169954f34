"""The two similarity consumers next to the matching path against the oracle's restatement of the
reference lines: MaskCut's normalised affinity + thresholding (evals/models/maskcut_processor.py:77-78, :103-106) and the
2AFC cosine prediction (evaluate_model_percepture.py:46-48, :118-122)."""
import importlib

import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bb():
    return importlib.import_module("midvision-probe_b200.backbones")


@pytest.mark.parametrize("C,N", [(768, 900), (768, 197), (384, 1024), (2048, 300), (64, 5)])
@pytest.mark.parametrize("dtype", ["f16", "tf32"])
def test_cosine_affinity_matches_oracle(mv, C, N, dtype):
    g = torch.Generator().manual_seed(C + N)
    base = torch.rand(C, 1, generator=g)                       # a common component: token features are nearly collinear
    feats = base + 0.4 * torch.randn(C, N, generator=g)
    want = restated.maskcut_affinity(feats)
    mv.correspondence.set_match_precision(dtype=dtype)
    try:
        got, val, idx = mv.affinity.cosine_affinity(feats, return_neighbours=True)
    finally:
        mv.correspondence.set_match_precision(dtype=mv.correspondence.DEFAULT_DTYPE)
    assert got.shape == (N, N) and got.dtype == torch.float32 and got.device.type == "cpu"
    tol = 1e-4 if dtype == "f16" else 1.5e-3                    # stated tolerances of the two operand types (tf32 truncates)
    assert (got - want).abs().max() <= tol, float((got - want).abs().max())
    assert (got.diagonal() - 1).abs().max() <= tol and (idx[:, 0] == torch.arange(N)).all()  # every token is its own best match
    torch.testing.assert_close(val[:, 0], got.max(dim=1).values, rtol=0, atol=0)              # the fused row maxima are the matrix's
    # thresholding: identical wherever the reference entry is clear of tau by the product tolerance
    tau = float(want.flatten().median())
    A, d = mv.affinity.threshold_affinity(got, tau, eps=1e-5)
    _, B, dB = restated.maskcut_affinity(feats, tau=tau, eps=1e-5)
    clear = (want - tau).abs() > tol
    assert torch.equal((A > 0.5)[clear], (B > 0.5)[clear]) and set(A.unique().tolist()) <= {1.0, float(torch.tensor(1e-5))}
    flips = (~clear).sum(dim=1).double()
    assert ((d - dB).abs() <= flips * (1 - 1e-5) + 1e-9).all()
    # and exactly the reference's own numbers when fed the reference's matrix
    A2, d2 = mv.affinity.threshold_affinity(want, tau, eps=1e-5)
    assert torch.equal(A2 > 0.5, B > 0.5) and torch.allclose(d2, dB, rtol=0, atol=1e-9)


def test_cosine_affinity_on_backbone_tokens(mv, bb):
    """MaskCut's real input: the last-block tokens of a ViT-B/16 (here random-init) at 480 x 480 -> 900 tokens."""
    model = bb.DenseViT(bb.vit_b16(0), multilayer=False).cuda()
    f = model(bb.smooth_images(5, 1, 480, 480).cuda())[0].float().cpu()      # (768, 30, 30)
    feats = f.reshape(768, -1)
    want = restated.maskcut_affinity(feats)
    got = mv.affinity.cosine_affinity(feats)
    assert (got - want).abs().max() <= 1e-4
    dev_out = mv.affinity.cosine_affinity(feats.cuda())
    assert dev_out.device.type == "cuda" and torch.equal(dev_out.cpu(), got)


@pytest.mark.parametrize("B,D", [(64, 768), (7, 2048), (300, 770), (1, 4)])
def test_twoafc_matches_oracle(mv, B, D):
    g = torch.Generator().manual_seed(B * D)
    ref = torch.randn(B, D, generator=g)
    left = ref + 0.8 * torch.randn(B, D, generator=g)
    right = ref + 0.8 * torch.randn(B, D, generator=g)
    left[0] = 0.0                                              # a zero vector: the eps clamp of cosine_similarity
    sl, sr, pred = mv.affinity.twoafc_predict(ref, left, right)
    wl, wr, wp = restated.twoafc(ref, left, right)
    torch.testing.assert_close(sl, wl, rtol=0, atol=2e-6)
    torch.testing.assert_close(sr, wr, rtol=0, atol=2e-6)
    clear = (wl - wr).abs() > 1e-5
    assert pred.dtype == torch.int64 and torch.equal(pred[clear], wp[clear])
