timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer" 2>&1 | tail -3
CASES=clustered:32:16384:16384,shifted:32:16384:16384,blob:32:16384:16384,blob:32:16384:1024,outliers:32:16384:16384,constant:8:16384:16384,uniform:32:16384:16384,sphere:32:16384:16384,planar:32:16384:16384
timeout 300 python tools/chamfer_algos.py --reps 10 --cases $CASES --json gpurun_out/r2_chamfer_algos_d.json 2>&1 | tail -9
for k in shifted:shifted uniform:blob clustered:clustered; do
k1=${k%%:*}; k2=${k##*:}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$k2.csv python tools/chamfer_step.py --steps 2 --no-backward --kind $k1 --kind2 $k2 > /dev/null 2>&1
done
