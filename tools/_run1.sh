timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer" 2>&1 | tail -5
python tools/chamfer_step.py --steps 4
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/r2_launches_b.csv python tools/chamfer_step.py --steps 5 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_b.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:60], r[-1])
PY
ncu --set full --clock-control none --import-source on -k regex:chamfer_dense_query -s 2 -c 1 -o gpurun_out/r2_dense_query_a python tools/chamfer_step.py --steps 4 --no-backward > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
timeout 300 python tools/chamfer_algos.py --json gpurun_out/r2_chamfer_algos_b.json --reps 10 2>&1 | tail -20
