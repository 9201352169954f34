"""Where the time of one host-tensor helper call goes (NAVI-shaped pair): H2D, graph replay, D2H + host work."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mv = importlib.import_module("midvision-probe_b200")
syn = importlib.import_module("midvision-probe_b200.synthetic")
C_ = mv.correspondence
pairs = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in syn.navi_pair(i).items()} for i in range(8)]
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


def load_replay(p):
    gm.load(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"])
    gm.graph.replay()
    torch.cuda.current_stream().synchronize()


def only_replay(p):
    gm.graph.replay()
    torch.cuda.current_stream().synchronize()


def only_load2(p):
    gm.load(p["feat_0"], p["feat_1"], p["xyz_grid_0"], p["xyz_grid_1"], two_streams=True)
    torch.cuda.current_stream().synchronize()


_streams = [torch.cuda.Stream() for _ in range(4)]


def only_load4(p):
    cur = torch.cuda.current_stream()
    C = p["feat_0"].shape[0]
    jobs = [(gm.f0[:C // 2], p["feat_0"][:C // 2]), (gm.f0[C // 2:], p["feat_0"][C // 2:]),
            (gm.f1[:C // 2], p["feat_1"][:C // 2]), (gm.f1[C // 2:], p["feat_1"][C // 2:])]
    for st, (dst, src) in zip(_streams, jobs):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            dst.copy_(src, non_blocking=True)
    gm.g0.copy_(p["xyz_grid_0"], non_blocking=True)
    gm.g1.copy_(p["xyz_grid_1"], non_blocking=True)
    for st in _streams:
        cur.wait_stream(st)
    cur.synchronize()


print(f"H2D of one pair (19.6 MB pinned) + sync : {timed(only_load):7.1f} us")
print(f"the same on four streams                : {timed(only_load4):7.1f} us")
print(f"the same on two streams                 : {timed(only_load2):7.1f} us")
print(f"graph replay + sync                     : {timed(only_replay):7.1f} us")
print(f"H2D + replay + sync                     : {timed(load_replay):7.1f} us")
print(f"full helper call                        : {timed(call):7.1f} us")
