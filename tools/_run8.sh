python -m pytest tests/test_gpu_k1.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -12 > gpurun_out/pytest_k1.log
for rows in split f32; do for kind in navi scannet; do
python tools/k1_probe.py --kind $kind --reps 3 --nosync --rows $rows 2>&1 | tail -1 > gpurun_out/k1p_${kind}_$rows.log
done; done
python bench.py --no-cpu-baseline --no-stress > gpurun_out/bench20.log 2>gpurun_out/bench20.err
cat gpurun_out/pytest_k1.log gpurun_out/k1p_*_split.log gpurun_out/k1p_*_f32.log; for f in gpurun_out/bench20.log; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pipe", d.get("e2e_pipeline"), "k2 TF", round(d["roofline"]["achieved"],1), d["recall"])
PY
done; tail -3 gpurun_out/bench20.err
