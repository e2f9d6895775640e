out=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/analysis/sanitize_cases.py > $out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -6 $out/sanitize_$tool.log
done
