for v in 0 1; do echo "--- ONEPASS=$v"; MVP_CHAMFER_BWD_ONEPASS=$v python tools/chamfer_step.py --steps 10 2>&1 | tail -3; done
MVP_CHAMFER_BWD_ONEPASS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "chamfer_backward or full_size" 2>&1 | tail -2
