"""Host-side logic that needs no GPU: pair sharding + the integer hit-count all-reduce over a 2-rank gloo
group, recall summaries, pure helpers, kernel-2 tile schedule arithmetic (restated in Python)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import restated

DEFAULT_DTYPE = "f16"  # correspondence.DEFAULT_DTYPE: what the finally blocks restore

THR3 = [0.01, 0.02, 0.05]
THR2 = [5, 25, 50]


def fake_pair_hits(i, n_counters):
    g = torch.Generator().manual_seed(i)
    h = torch.randint(0, 1000, (n_counters,), generator=g, dtype=torch.int64)
    h[0] = 1000
    return h


def _worker(rank, world, port, num_pairs, out_dir):
    import importlib

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = importlib.import_module("midvision-probe_b200.evaluation")
    acc = ev.RecallAccumulator(THR3, THR2, device="cpu")
    mine = list(ev.shard_pairs(num_pairs, rank, world))
    for i in mine:
        acc.merge_(fake_pair_hits(i, acc.hits.numel()))
    acc.all_reduce()
    torch.save({"hits": acc.hits, "mine": mine, "summary": acc.summary()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_pairs", [7, 10])
def test_two_rank_counts_equal_single_process(tmp_path, mv, num_pairs):
    world = 2
    port = 29500 + (os.getpid() % 500) + num_pairs
    mp.spawn(_worker, args=(world, port, num_pairs, str(tmp_path)), nprocs=world, join=True)
    ev = mv.evaluation
    single = ev.RecallAccumulator(THR3, THR2, device="cpu")
    for i in range(num_pairs):
        single.merge_(fake_pair_hits(i, single.hits.numel()))
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert sorted(outs[0]["mine"] + outs[1]["mine"]) == list(range(num_pairs))      # a partition of the pairs
    assert not set(outs[0]["mine"]) & set(outs[1]["mine"])
    for o in outs:
        assert torch.equal(o["hits"], single.hits)                                    # identical integers on every rank
        assert o["summary"] == single.summary()


def test_recall_summary_matches_float_mean(mv):
    ev = mv.evaluation
    acc = ev.RecallAccumulator(THR3, THR2, device="cpu")
    g = torch.Generator().manual_seed(0)
    e3 = torch.rand(4000, generator=g) * 0.08
    e2 = torch.rand(4000, generator=g) * 60
    mu = torch.rand(4000, generator=g) < 0.3
    h = [4000, int(mu.sum())] + [int((e3 < t).sum()) for t in THR3] + [int((e2 < t).sum()) for t in THR2]
    h += [int(((e3 < t) & mu).sum()) for t in THR3] + [int(((e2 < t) & mu).sum()) for t in THR2]
    acc.merge_(torch.tensor(h))
    s = acc.summary()
    for t, r in zip(THR3, restated.recall(e3, THR3)):
        assert abs(s["recall_3d"][t] - r) < 1e-4
    for t, r in zip(THR2, restated.recall(e2, THR2)):
        assert abs(s["recall_2d"][float(t)] - r) < 1e-4
    assert abs(s["mutual_recall_3d"][THR3[0]] - 100.0 * ((e3 < THR3[0]) & mu).sum().item() / mu.sum().item()) < 1e-9
    with pytest.raises(ValueError):
        ev.RecallAccumulator(list(range(17)), [], device="cpu")


def test_pure_helpers_match_oracle(mv, golden):
    C_ = mv.correspondence
    g = golden("rows_small")
    torch.testing.assert_close(C_.get_grid(3, 5), torch.from_numpy(g["grid"]), rtol=0, atol=0)
    d = torch.from_numpy(g["dists"])
    torch.testing.assert_close(C_.calculate_ratio_test(d), torch.from_numpy(g["ratio"]), rtol=0, atol=0)
    xyz = torch.randn(11, 3) + torch.tensor([0.0, 0.0, 3.0])
    Kmat = torch.tensor([[500.0, 0, 320], [0, 500, 240], [0, 0, 1]])
    torch.testing.assert_close(C_.project_3dto2d(xyz, Kmat), restated.project_3dto2d(xyz, Kmat), rtol=0, atol=0)
    y = torch.arange(10.0)
    x = torch.tensor([5.0, 35, 65, 95, 10, 40, 70, 100, 119, 121])
    b = C_.compute_binned_performance(y, x, [0, 30, 60, 90, 120])
    assert [float(v) for v in b] == [2.0, 3.0, 4.0, 6.0]
    errs = [0.5, 1.5, 2.5, 7.0]
    auc = C_.error_auc(errs, [5, 10])
    assert len(auc) == 2 and 0 < auc[0] < 1 and auc[1] > auc[0]


def test_feature_layout_and_dtype_detection(mv):
    """which feature tensors take the zero-copy / 16-bit hand-off paths is decided from strides and dtype alone."""
    C_ = mv.correspondence
    chw = torch.zeros(8, 3, 5)
    hwc_view = torch.zeros(3, 5, 8).permute(2, 0, 1)          # what a (B, tokens, C) ViT output looks like as (C, h, w)
    assert not C_._is_channel_last(chw) and C_._is_channel_last(hwc_view)
    assert C_._is_channel_last(hwc_view.to(torch.bfloat16)) and C_._is_channel_last(hwc_view.half())
    assert not C_._is_channel_last(hwc_view.double())          # fp64 goes through the fp32 conversion
    assert not C_._is_channel_last(torch.zeros(8, 1, 1))       # degenerate strides: ambiguous, take the copy path
    assert not C_._is_channel_last(torch.zeros(3, 5, 16)[:, :, ::2].permute(2, 0, 1))  # strided channels
    assert C_._row_format("split") == (True, False, True) and C_._row_format("f32") == (True, True, False)
    C_.set_match_precision(dtype="tf32")
    try:
        assert C_._row_format("split") == (False, True, False)  # the tf32 path always keeps fp32 rows
    finally:
        C_.set_match_precision(dtype=DEFAULT_DTYPE)


def test_argument_errors_mirror_the_reference(mv):
    C_ = mv.correspondence
    with pytest.raises(AssertionError):  # the reference's `assert metric in [...]` (correspondence.py:45)
        C_.knn_points(torch.zeros(2, 8), torch.zeros(2, 8), 1, "manhattan")
    with pytest.raises(AssertionError):
        C_.get_correspondences_ratio_test(torch.zeros(2, 8), torch.zeros(2, 8), 1, metric="manhattan")
    with pytest.raises(ValueError):
        C_.set_match_precision(dtype="fp8")
    with pytest.raises(ValueError):
        C_.set_match_precision(rows="fp8")


# ---- kernel 2's tile schedule, restated: every tile owned exactly once, parts consecutive -----------------
def sched(n, m, mc, clusters):
    n_sb = -(-n // (128 * mc))
    n_ct = -(-m // 256)
    T = n_sb * n_ct
    G = max(1, min(clusters, T))
    return n_sb, n_ct, T, G


@pytest.mark.parametrize("n,m,mc,clusters", [(19200, 19200, 1, 148), (19200, 19200, 2, 74), (300, 280, 1, 148), (1, 1, 1, 148),
                                             (5025, 5025, 1, 148), (12544, 12544, 4, 33), (1 << 20, 300, 1, 148)])
def test_k2_schedule_partitions_tiles_and_slots_are_unique(n, m, mc, clusters):
    n_sb, n_ct, T, G = sched(n, m, mc, clusters)
    begin = lambda c: c * T // G
    owner = lambda t: ((t + 1) * G + T - 1) // T - 1
    assert begin(0) == 0 and begin(G) == T
    slots = set()
    step = max(1, T // 5000)
    for c in range(G):
        assert begin(c + 1) > begin(c)                      # every cluster below G owns at least one tile
        for t in {begin(c), begin(c + 1) - 1}:
            assert owner(t) == c
        sbs = range(begin(c) // n_ct, (begin(c + 1) - 1) // n_ct + 1)
        for sb in sbs:
            assert (c + sb) not in slots                    # partial-record slot = cluster + row block: unique
            slots.add(c + sb)
    assert max(slots) < clusters + n_sb                     # fits the workspace mv_k2_workspace_bytes sizes
    for sb in range(0, n_sb, max(1, n_sb // 50)):
        first, last = owner(sb * n_ct), owner((sb + 1) * n_ct - 1)
        assert all((c + sb) in slots for c in range(first, last + 1))


def test_bench_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--gpus", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""    # other ranks exit 0 without work
