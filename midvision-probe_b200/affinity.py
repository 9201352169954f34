"""The two similarity consumers next to the matching path, on the same kernels.

    cosine_affinity(feats)                 MaskCut's normalised affinity  F.normalize(feats, dim=0); feats^T @ feats
                                           (evals/models/maskcut_processor.py:77-78): kernel 1 (normalise + f16c rows)
                                           -> kernel 2 with the similarity tiles written out (mv_k2_affinity)
    threshold_affinity(S, tau, eps)        A = S > tau ? 1 : eps and the degrees d_i = sum_j A_ij  (:103-106)
    twoafc_predict(ref, left, right)       the 2AFC evaluation's two cosine similarities per sample and the
                                           prediction (evaluate_model_percepture.py:46-48, :118-122)

The k-means threshold search (sklearn, :80-96), the eigen-decomposition and the metric bookkeeping stay where they are.
"""
from ctypes import c_float, c_size_t

import torch

from . import _lib as L
from . import correspondence as C_

__all__ = ["cosine_affinity", "threshold_affinity", "twoafc_predict"]


def cosine_affinity(feats, return_neighbours=False):
    """(N, N) fp32 cosine affinity of the N columns of feats (C, N) -- the `A` of get_affinity_matrix before its
    `.cpu().numpy()` (maskcut_processor.py:77-78) -- on the device of `feats`.

    The product runs on kernel 2 (tcgen05) with the operand type of set_match_precision.  Stated tolerance against the
    fp32 product: 1e-4 with the default f16c rows (measured <= 6e-5; the closer to collinear the token features are, the
    smaller -- a plain bf16 product is off by up to 1e-3), 1.5e-3 with "tf32" (the tensor core truncates fp32 operands).  return_neighbours: also the per-row top-2 (values, indices) that
    the same launch produces."""
    dev = C_._device()
    in_dev = feats.device
    X = C_._f32(feats, dev).t().contiguous()  # (N, C) rows = tokens
    n, C = X.shape
    C_._check_C(C)
    if n == 0:
        return torch.zeros((0, 0), dtype=torch.float32, device=in_dev)
    (A16, A32), (B16, B32), mu = C_._rows_pair(X, X, dev)
    tf32 = C_._CFG["dtype"] == "tf32"
    f16 = C_._CFG["dtype"] == "f16"
    A, B = (A32, B32) if tf32 else (A16, B16)
    ld_s = (n + 3) // 4 * 4
    S = torch.empty((n, ld_s), dtype=torch.float32, device=dev)
    row_val = torch.empty((n, 2), dtype=torch.float32, device=dev)
    row_idx = torch.empty((n, 2), dtype=torch.int32, device=dev)
    col_best = torch.empty((n,), dtype=torch.int64, device=dev)
    wsb = L.load().mv_k2_workspace_bytes(n, n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    L.call("mv_k2_affinity", L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], n, n, C + 8 if f16 else C, None, None,
           L.MV_DTYPE_TF32 if tf32 else (L.MV_DTYPE_F16 if f16 else L.MV_DTYPE_BF16), C_._CFG["cluster"], L.ptr(S), ld_s,
           L.ptr(row_val), L.ptr(row_idx), L.ptr(col_best), L.ptr(ws), c_size_t(wsb), C_._stream())
    out = S[:, :n]
    out = out.contiguous() if ld_s != n else out
    if return_neighbours:
        return out.to(in_dev), row_val.to(in_dev), row_idx.long().to(in_dev)
    return out.to(in_dev)


def threshold_affinity(S, tau, eps=1e-5):
    """(A, d): A = where(S > tau, 1, eps) as fp32 and the degrees d_i = sum_j A_ij as fp64, computed from the integer
    counts of entries above tau (maskcut_processor.py:103-106; D = diag(d))."""
    dev = C_._device()
    in_dev = S.device
    Sd = C_._f32(S, dev)
    n, m = Sd.shape
    A = torch.empty((n, m), dtype=torch.float32, device=dev)
    cnt = torch.empty((n,), dtype=torch.int32, device=dev)
    L.call("mv_affinity_threshold", L.ptr(Sd), n, m, Sd.stride(0), c_float(float(tau)), c_float(float(eps)), L.ptr(A), L.ptr(cnt),
           C_._stream())
    d = cnt.double() + (m - cnt.double()) * float(eps)
    return A.to(in_dev), d.to(in_dev)


def twoafc_predict(features_ref, features_left, features_right):
    """(similarity_left, similarity_right, predictions): F.cosine_similarity(ref, left / right, dim=-1) per sample and
    torch.where(similarity_left > similarity_right, 0, 1) (evaluate_model_percepture.py:46-48, :118-122), one launch."""
    dev = C_._device()
    in_dev = features_ref.device
    r, a, b = (C_._f32(t, dev) for t in (features_ref, features_left, features_right))
    if r.shape != a.shape or r.shape != b.shape or r.dim() != 2:
        raise ValueError("features must be three (B, D) tensors of the same shape")
    n, D = r.shape
    if D % 4:
        pad = 4 - D % 4  # zero columns change neither the dot products nor the norms
        r, a, b = (torch.nn.functional.pad(t, (0, pad)).contiguous() for t in (r, a, b))
        D += pad
    sl = torch.empty((n,), dtype=torch.float32, device=dev)
    sr = torch.empty((n,), dtype=torch.float32, device=dev)
    pred = torch.empty((n,), dtype=torch.int32, device=dev)
    if n:
        L.call("mv_cosine_2afc", L.ptr(r), L.ptr(a), L.ptr(b), n, D, L.ptr(sl), L.ptr(sr), L.ptr(pred), C_._stream())
    return sl.to(in_dev), sr.to(in_dev), pred.long().to(in_dev)
