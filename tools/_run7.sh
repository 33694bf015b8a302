timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer" 2>&1 | tail -2
for f in 0 1; do echo "--- MVP_GRID_DYNROWS=$f"; MVP_GRID_DYNROWS=$f timeout 300 python tools/chamfer_algos.py --reps 20 --cases uniform:32:16384:16384,sphere:32:16384:16384,planar:32:16384:16384,uniform:64:2048:2048,uniform:64:2048:3072,uniform:32:16384:1024 2>&1 | tail -6; done
ncu --set full --clock-control none --import-source on -k regex:chamfer_grid_query -s 1 -c 1 -o gpurun_out/r2_query_dyn -f python tools/chamfer_step.py --steps 2 --no-backward > /dev/null 2>&1
