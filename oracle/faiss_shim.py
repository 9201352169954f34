"""TEST INFRASTRUCTURE ONLY.  Stand-in for the four faiss-gpu symbols the reference touches
(evals/utils/correspondence.py:4-5, :11, :20-22): StandardGpuResources, GpuIndexFlatL2(res, d) with
.add(x) / .search(q, k), and the faiss.contrib.torch_utils import.

faiss-gpu 1.8.0 GpuIndexFlatL2 is an exact (non-approximate) brute-force index: squared L2 distances
||q||^2 - 2 q.x + ||x||^2 from an fp32 GEMM, k smallest per query in ascending order, int64 labels.
This shim computes the same thing in fp32 torch on the CPU.
"""
import sys
import types

import torch


class StandardGpuResources:
    pass


class GpuIndexFlatL2:
    def __init__(self, res, d):
        self.d = d
        self.x = None

    def add(self, x):
        assert x.shape[1] == self.d
        self.x = x.detach().float().cpu()

    def search(self, q, k):
        q = q.detach().float().cpu()
        d2 = (q * q).sum(1, keepdim=True) - 2.0 * (q @ self.x.t()) + (self.x * self.x).sum(1)[None, :]
        dist, idx = torch.topk(d2, k, dim=1, largest=False, sorted=True)
        return dist, idx.long()


def install():
    """Register the stand-in as `faiss` / `faiss.contrib.torch_utils` in sys.modules."""
    if "faiss" in sys.modules and not getattr(sys.modules["faiss"], "_mv_shim", False):
        return sys.modules["faiss"]  # a real faiss is present: use it
    faiss = types.ModuleType("faiss")
    faiss._mv_shim = True
    faiss.StandardGpuResources = StandardGpuResources
    faiss.GpuIndexFlatL2 = GpuIndexFlatL2
    contrib = types.ModuleType("faiss.contrib")
    tu = types.ModuleType("faiss.contrib.torch_utils")
    contrib.torch_utils = tu
    faiss.contrib = contrib
    sys.modules["faiss"] = faiss
    sys.modules["faiss.contrib"] = contrib
    sys.modules["faiss.contrib.torch_utils"] = tu
    return faiss
