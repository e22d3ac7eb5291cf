#!/bin/bash
# compute-sanitizer over tools/sanitize.py (every kernel kind at small sizes): memcheck, synccheck, racecheck
mkdir -p gpurun_out
python tools/sanitize.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -2 gpurun_out/sanitize_plain.log
for tool in memcheck synccheck racecheck; do
  start=$(date +%s)
  timeout ${SAN_TIMEOUT:-1200} compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 python tools/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  echo "== $tool rc=$rc $(( $(date +%s) - start )) s"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize.py: done|Invalid|hazard|Error" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -12
done
