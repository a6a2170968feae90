#!/bin/bash
# compute-sanitizer over every kernel variant on small shapes (GPU box).  Output: gpurun_out/sanitizer_<round>.txt
R=${1:-r01}
OUT=gpurun_out/sanitizer_${R}.txt
: > $OUT
for tool in memcheck racecheck; do
  for args in "cfg1 1 1 fast" "cfg1 1 1 generic" "cfg2 1 1 fast" "cfg2 1 1 generic"; do
    echo "=== compute-sanitizer --tool $tool  run_once.py $args (small map)" >> $OUT
    SRCV_SMALL=1 timeout 250 compute-sanitizer --tool $tool --print-limit 5 python scripts/run_once.py $args 2>&1 \
      | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race|hazard|^cfg" | head -8 >> $OUT
  done
done
cat $OUT
