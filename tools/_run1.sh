timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer" 2>&1 | tail -3
timeout 300 python tools/chamfer_algos.py --json gpurun_out/r2_chamfer_algos_b.json --reps 10 --cases clustered:32:16384:16384,shifted:32:16384:16384,constant:8:16384:16384,blob:32:16384:16384,blob:32:16384:1024,uniform:32:16384:16384 2>&1 | tail -20
ncu --set full --clock-control none --import-source on -k regex:chamfer_rest -s 0 -c 1 -o gpurun_out/r2_rest_blob -f python tools/chamfer_step.py --steps 1 --no-backward --kind uniform --kind2 blob > /dev/null 2>&1
