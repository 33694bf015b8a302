O=gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_ops.py --small > $O/r2_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_ops:' $O/r2_sanitize_$tool.log | tr '\n' ' ')"
done
