python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
for rows in split f32; do
python tools/k1_probe.py --kind navi --reps 3 --nosync --rows $rows 2>&1 | tail -1 > gpurun_out/k1p_navi_$rows.log
python tools/k1_probe.py --kind scannet --reps 3 --nosync --rows $rows 2>&1 | tail -1 > gpurun_out/k1p_scannet_$rows.log
MVMATCH_ROWS=$rows python bench.py --no-cpu-baseline --no-stress > gpurun_out/bench19_$rows.log 2>gpurun_out/bench19_$rows.err
done
MVMATCH_ROWS=split python bench.py --no-cpu-baseline --no-stress --workload scannet > gpurun_out/bench19_scannet_split.log 2>gpurun_out/bench19_scannet_split.err
cat gpurun_out/pytest_gpu.log gpurun_out/k1p_*_split.log gpurun_out/k1p_*_f32.log; for f in gpurun_out/bench19*.log; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "k2 TF", round(d["roofline"]["achieved"],1), d["recall"])
PY
done
