#!/bin/bash
# compute-sanitizer passes of round 2 (racecheck / synccheck / memcheck); summaries -> gpurun_out/r2/sanitize_*.txt
OUT=gpurun_out/r2
mkdir -p $OUT
for tool in ${SAN_TOOLS:-racecheck synccheck memcheck}; do
  for what in ${SAN_WHAT:-fsst lstm overlap pipeline train}; do
    echo "=== $tool $what ($(date +%T))"
    timeout -s KILL ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_r2.py $what > $OUT/sanitize_${tool}_${what}.txt 2>&1
    echo "rc=$?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|^(fsst|lstm|overlap|pipeline|train) " $OUT/sanitize_${tool}_${what}.txt | head -12
  done
done
