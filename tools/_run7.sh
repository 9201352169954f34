python tools/h2d_probe.py > gpurun_out/h2d.log 2>&1
python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
cat gpurun_out/h2d.log gpurun_out/e2e_probe.log
