"""Where the time of one PAGEABLE host-tensor helper call goes (NAVI-shaped pair)."""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
C_ = mv.correspondence
pairs = [syn.navi_pair(i) for i in range(8)]
call = lambda p: C_.estimate_correspondence_xyz(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], 1000)
for p in pairs[:3]:
    call(p)
gm = next(iter(C_._HELPER_GRAPHS.values()))


def timed(fn, n=40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(pairs[i % 8])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def only_load(p):
    gm.load(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"])
    torch.cuda.current_stream().synchronize()


def load_feats(p):
    mv.evaluation.copy_in(gm.f0, p["feat_0"])
    mv.evaluation.copy_in(gm.f1, p["feat_1"])
    torch.cuda.current_stream().synchronize()


def load_grids(p):
    mv.evaluation.copy_in(gm.g0, p["xyz_grid_0"])
    mv.evaluation.copy_in(gm.g1, p["xyz_grid_1"])
    torch.cuda.current_stream().synchronize()


def only_replay(p):
    gm.graph.replay()
    torch.cuda.current_stream().synchronize()


print(f"threads {mv.load().mv_h2d_staged_threads()}: call {timed(call):.0f} us | load {timed(only_load):.0f} (feats {timed(load_feats):.0f}, grids {timed(load_grids):.0f}) | "
      f"replay {timed(only_replay):.0f}")
