python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -8 > gpurun_out/pytest_pipe.log
python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
cat gpurun_out/pytest_pipe.log gpurun_out/e2e_probe.log
