python -m pytest tests/test_gpu_k1.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_k1.log
for kind in navi scannet; do
  python tools/k1_probe.py --kind $kind --reps 5 --nosync 2>&1 | tail -1 > gpurun_out/k1p_${kind}_new.log
done
python - > gpurun_out/memset.log 2>&1 <<'PY'
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
print(f"memset 1 GiB: {10 * (1 << 30) / e0.elapsed_time(e1) / 1e6:.0f} GB/s write-only")
y = torch.empty(200 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): y.zero_()
torch.cuda.synchronize()
e0.record()
for _ in range(10): y.zero_()
e1.record(); torch.cuda.synchronize()
print(f"memset 200 MiB: {10 * (200 << 20) / e0.elapsed_time(e1) / 1e6:.0f} GB/s write-only")
PY
cat gpurun_out/pytest_k1.log gpurun_out/k1p_*_new.log gpurun_out/memset.log
