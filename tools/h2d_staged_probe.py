"""Upload rate of mv_h2d_staged (pageable -> device through the threaded pinned ring) against torch's own pageable and
pinned copies.  MVMATCH_STAGE_THREADS is read once per process: run once per thread count."""
import importlib
import os
import sys
import time
from ctypes import c_void_p

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mv = importlib.import_module("midvision-probe_b200")
L = mv._lib
nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
src = torch.randn(nbytes // 4)
pin = src.clone().pin_memory()
dst = torch.empty(nbytes // 4, device="cuda")
st = torch.cuda.current_stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


staged = timed(lambda: L.call("mv_h2d_staged", c_void_p(dst.data_ptr()), c_void_p(src.data_ptr()), nbytes, c_void_p(st.cuda_stream)))
assert torch.equal(dst.cpu(), src)
t0 = time.perf_counter()
tmp = src.clone()
host_copy = nbytes / (time.perf_counter() - t0) / 1e9
print(f"threads {L.load().mv_h2d_staged_threads()} cpus {len(os.sched_getaffinity(0))}: staged {staged:.1f} GB/s | torch pageable "
      f"{timed(lambda: dst.copy_(src, non_blocking=True)):.1f} | torch pinned {timed(lambda: dst.copy_(pin, non_blocking=True)):.1f} | "
      f"one-thread host memcpy {host_copy:.1f} GB/s")
