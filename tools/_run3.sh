for k in shifted:shifted uniform:blob clustered:clustered outliers:outliers; do
k1=${k%%:*}; k2=${k##*:}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$k2.csv python tools/chamfer_step.py --steps 2 --no-backward --kind $k1 --kind2 $k2 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:chamfer_rest_kernel -s 1 -c 1 -o gpurun_out/r2_rest_shifted -f python tools/chamfer_step.py --steps 2 --no-backward --kind shifted > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chamfer_rest_kernel -s 1 -c 1 -o gpurun_out/r2_rest_blob2 -f python tools/chamfer_step.py --steps 2 --no-backward --kind uniform --kind2 blob > /dev/null 2>&1
