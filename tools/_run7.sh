ncu --set full --clock-control none --import-source on -k regex:"chamfer_grid_query|chamfer_rest_kernel" -s 2 -c 2 -o gpurun_out/r2_blob_qr -f python tools/chamfer_step.py --steps 2 --no-backward --kind uniform --kind2 blob > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r2_blob_qr.ncu-rep | grep -E "^### |time_duration|thread_inst|inst_executed.sum|issue_active.avg.pct_of_peak_sustained_active|long_scoreboard|lg_throttle|l1tex__throughput|lts__throughput|warps_active"
python tools/ncu_regions.py gpurun_out/r2_blob_qr.ncu-rep | head -30
