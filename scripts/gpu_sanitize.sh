#!/bin/bash
# compute-sanitizer passes over small ragged invocations of every kernel family + the new tests
set +e
TAG=${1:-r02}
mkdir -p gpurun_out/$TAG
O=gpurun_out/$TAG
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$|MISMATCH|Error|hazard" $O/sanitizer_$tool.log | sort | uniq -c | head -20
done
