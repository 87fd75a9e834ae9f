for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize.py 2>&1 | grep -E "ok |ERROR SUMMARY|RACECHECK SUMMARY|Error|error:|hazard" | head -12
done
