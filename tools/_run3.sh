python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench15.log 2>gpurun_out/bench15.err
cat gpurun_out/pytest_gpu.log; tail -c 3000 gpurun_out/bench15.log; tail -5 gpurun_out/bench15.err
