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


def test_switch_entry_points_without_a_gpu(mv):
    """the setters that only flip host-side state: previous value returned, out-of-range values only query."""
    lib = mv.load()
    assert lib.mv_spair_set_heatmap_terms(-1) == 3          # default: 3xTF32
    assert lib.mv_spair_set_heatmap_terms(1) == 3
    assert lib.mv_spair_set_heatmap_terms(2) == 1           # not a mode: query only
    assert lib.mv_spair_set_heatmap_terms(3) == 1
    assert mv.spair.set_heatmap_precision("3xtf32") == "3xtf32"
    with pytest.raises(KeyError):
        mv.spair.set_heatmap_precision("fp8")
    prev = lib.mv_k2_set_streamk(-1)
    assert prev in (0, 1) and lib.mv_k2_set_streamk(-1) == prev


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


def test_every_compute_entry_point_validates_before_touching_cuda(mv):
    """argument errors are reported through the status code + mv_last_error() before any CUDA call, so they can be
    exercised without a GPU: null pointers, non-positive sizes, misaligned buffers, out-of-range sizes."""
    import ctypes

    lib = mv.load()
    P16, P4 = ctypes.c_void_p(16), ctypes.c_void_p(4)
    E_ARG, E_ALIGN, E_RANGE = -1, -2, -3
    f = ctypes.c_float
    cases = [
        ("mv_chw_to_hwc", (None, P16, 8, 4, 0, None, None), E_ARG),
        ("mv_chw_to_hwc", (P16, P16, 8, 4, 1, None, None), E_ARG),                     # prenorm without scratch
        ("mv_feat_to_hwc_f32", (P16, 9, 0, 8, 4, P16, None), E_ARG),                   # unknown dtype
        ("mv_compact_valid", (P16, 1, (1 << 20) + 1, P16, P16, None), E_RANGE),
        ("mv_geom_backproject", (P16, 0, 4, P16, P16, None), E_ARG),
        ("mv_geom_project_coords", (None, None, None, 4, P16, 4, 4, 2, 2, P16, P16, None), E_ARG),
        ("mv_geom_grid_coords", (P16, None, None, 4, 4, 4, 0, 2, P16, P16, P16, None), E_ARG),
        ("mv_geom_keypoint_coords", (P16, 1, 4, f(224.0), 2, 2, P16, None), E_ARG),   # stride < 2
        ("mv_k1_sample_normalize", (0, P16, 12, 2, 2, P16, None, 4, 1, P16, None, None, None, None), E_RANGE),  # C % 8
        ("mv_k1_sample_normalize", (0, P4, 8, 2, 2, P16, None, 4, 1, P16, None, None, None, None), E_ALIGN),
        ("mv_k1_sample_normalize", (0, P16, 8, 2, 2, P16, None, 4, 1, None, P16, None, None, None), E_ARG),     # lo without hi
        ("mv_k2_sim_top2", (None, P16, 4, 4, 8, None, None, 0, 0, P16, P16, P16, P16, ctypes.c_size_t(1 << 20), None), E_ARG),
        ("mv_k2_sim_top2", (P16, P16, 4, 4, 8, None, None, 7, 0, P16, P16, P16, P16, ctypes.c_size_t(1 << 20), None), E_ARG),
        ("mv_k3_ratio_mutual", (P16, P16, 6, None, 4, P16, None, 1, None, None, None, None), E_ALIGN),          # C % 4
        ("mv_k3_ratio_mutual_split", (P16, P16, P16, None, 8, None, 4, P16, None, 1, None, None, None, None), E_ARG),
        ("mv_k3_ratio_mutual_split", (P16, P16, P16, P16, 12, None, 4, P16, None, 1, None, None, None, None), E_ALIGN),
        ("mv_gather_rows", (P16, 0, P16, None, 4, P16, None), E_ARG),
        ("mv_pack_matches", (P16, P16, P16, None, 4, P16, P16, P16, None, None, None, P16, None), E_ARG),       # uv0 without uv1
        ("mv_argmax_rows", (P16, 2, 0, 1, P16, None), E_ARG),
        ("mv_k3_spair_errors", (P16, 65, 14, P16, P16, 3, f(224.0), f(1.0), f(0.1), None, None, None, None, None, None, 0, None), E_RANGE),
        ("mv_spair_match_batch", (P16, 2, 8, 4, 4, P16, P16, 65, 3, P16, f(224.0), f(0.1), None, None, None, None, None, None, 0, None), E_RANGE),
        ("mv_spair_match_batch", (None, 2, 8, 4, 4, P16, P16, 4, 3, P16, f(224.0), f(0.1), None, None, None, None, None, None, 0, None), E_ARG),
        ("mv_spair_match_batch", (P16, 2, 8, 4, 4, P16, P16, 4, 3, P16, f(224.0), f(0.1), None, None, None, None, None, P16, 2, None), E_ARG),
    ]
    for name, args, want in cases:
        rc = getattr(lib, name)(*args)
        assert rc == want, f"{name}{args}: rc={rc}, wanted {want}: {lib.mv_last_error()}"
        assert name.encode() in lib.mv_last_error()
    # empty work is not an error
    assert lib.mv_spair_match_batch(None, 0, 8, 4, 4, None, None, 4, 3, None, f(224.0), f(0.1), None, None, None, None, None, None, 0, None) == 0
    assert lib.mv_k1_sample_normalize(2, P16, 8, 0, 0, None, None, 0, 1, P16, None, None, None, None) == 0
