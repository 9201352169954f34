"""The C-ABI library loads and exports exactly what include/mvmatch.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "mvmatch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(mv):
    lib = mv.load()
    names = header_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in mvmatch.h but not exported by libmvmatch.so"


def test_python_prototypes_cover_the_header(mv):
    assert sorted(mv._lib.PROTOTYPES) == header_functions()


def test_no_torch_symbols_in_the_abi(mv):
    import subprocess

    out = subprocess.run(["nm", "-D", "--undefined-only", mv._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "at::" not in out and "c10" not in out and "torch" not in out


def test_version_and_argument_errors_without_a_gpu(mv):
    lib = mv.load()
    assert lib.mv_version() == 100
    # argument validation happens before any CUDA call
    rc = lib.mv_k1_sample_normalize(7, None, 8, 1, 1, None, None, 1, 0, None, None, None, None, None)
    assert rc == -1
    assert b"mv_k1_sample_normalize" in lib.mv_last_error()
    rc = lib.mv_k3_topk_matches(ctypes.c_void_p(16), ctypes.c_void_p(16), None, 10, 1 << 20, ctypes.c_void_p(16),
                                ctypes.c_void_p(16), ctypes.c_void_p(16), None, None)
    assert rc == -3
    with pytest.raises(mv.MvMatchError):
        mv._lib.call("mv_compact_valid", None, 1, 4, None, None, None)


def test_product_has_no_oracle_import():
    pkg = os.path.join(ROOT, "midvision-probe_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), f"{f} must not import oracle/"
            assert "oracle/" not in src and "oracle." not in src, f"{f} must not reference oracle/"


def test_product_raises_without_cuda(mv):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        mv.correspondence.argmax_2d(torch.zeros(2, 3, 4))
