for kind in navi scannet; do
ncu --set full --clock-control none --import-source on -k regex:k1_rows -s 2 -c 1 -f -o gpurun_out/k1t_${kind} python tools/k1_probe.py --kind $kind --reps 2 --nosync > gpurun_out/ncu_k1t_${kind}.log 2>&1
done
ls -la gpurun_out/k1t_*
