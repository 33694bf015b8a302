set -u
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"chamfer_grid|chamfer_rest|chamfer_grad" -s 6 -c 6 -o $O/r2_chamfer_step -f python tools/chamfer_step.py --steps 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"fps|gather|group|scatter|interp|three_nn|topk|knn|ball|chamfer_grid" -o /tmp/r2_ops -f python tools/profile_ops.py fps gather group three_nn knn_points ball_query knn > $O/r2_ops.log 2>&1
python tools/ncu_summary.py full /tmp/r2_ops.ncu-rep > $O/r2_ops_full.md
ncu --set full --clock-control none -k regex:"chamfer_pair|chamfer_resolve|emd" -o /tmp/r2_ops2 -f python tools/profile_ops.py chamfer_brute emd > $O/r2_ops2.log 2>&1
python tools/ncu_summary.py full /tmp/r2_ops2.ncu-rep > $O/r2_ops2_full.md
ls -la $O | tail -6
