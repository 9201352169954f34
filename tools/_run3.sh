python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --no-stress > gpurun_out/bench16.log 2>gpurun_out/bench16.err
python bench.py --no-cpu-baseline --no-stress --feat-layout hwc > gpurun_out/bench16_hwc.log 2>gpurun_out/bench16_hwc.err
python bench.py --no-cpu-baseline --no-stress --workload scannet > gpurun_out/bench16_scannet.log 2>gpurun_out/bench16_scannet.err
cat gpurun_out/pytest_gpu.log; for f in gpurun_out/bench16*.log; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "k2 TF", round(d["roofline"]["achieved"],1), d["recall"]["recall_3d"])
PY
done; tail -3 gpurun_out/bench16*.err
