python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py --workload spair > gpurun_out/bench17_spair.log 2>gpurun_out/bench17_spair.err
for kind in navi scannet; do
ncu --set full --clock-control none --import-source on -k regex:k1_warp_rows -s 2 -c 1 -f -o gpurun_out/k1w2_${kind} python tools/k1_probe.py --kind $kind --reps 2 --nosync > gpurun_out/ncu_k1w2_${kind}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:spair_batch -s 1 -c 1 -f -o gpurun_out/spair_batch python tools/spair_probe.py --pairs 592 > gpurun_out/ncu_spair.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stress > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench18.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stress > gpurun_out/ncu_bench18.log 2>&1
cat gpurun_out/pytest_gpu.log; cat gpurun_out/bench17_spair.log | cut -c1-1500; python tools/launch_summary.py gpurun_out/launches_bench18.csv | head -20
