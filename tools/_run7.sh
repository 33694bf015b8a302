timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer or three_nn or knn_points" 2>&1 | tail -2
timeout 300 python tools/chamfer_algos.py --reps 30 --cases uniform:32:16384:16384,sphere:32:16384:16384,planar:32:16384:16384,uniform:64:2048:2048,uniform:64:2048:3072,uniform:32:16384:1024 2>&1 | tail -6 | cut -c1-130
ncu --set full --clock-control none --import-source on -k regex:chamfer_grid_query -s 1 -c 1 -o gpurun_out/r2_query_pool -f python tools/chamfer_step.py --steps 2 --no-backward > /dev/null 2>&1
