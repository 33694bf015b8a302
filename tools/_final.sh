set -u
O=gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; python -c "
import json; d=json.loads(open('$O/r2_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['frac_of_copy_ceiling'], d['direct_launch']['ms_per_step'], d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chamfer_grid|chamfer_rest|chamfer_grad" -s 5 -c 5 -o $O/r2_chamfer_step -f python tools/chamfer_step.py --steps 2 > /dev/null 2>&1
python tools/ncu_summary.py full $O/r2_chamfer_step.ncu-rep | grep -E "^### |time_duration"
