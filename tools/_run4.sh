python -m pytest tests/test_gpu_k3.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_k3.log
python tools/spair_probe.py > gpurun_out/spair_probe.log 2>&1
cat gpurun_out/pytest_k3.log gpurun_out/spair_probe.log
