#!/bin/bash
# compute-sanitizer passes over scripts/sanitize_small.py (run on the GPU box); summary -> gpurun_out/sanitizer.txt
out=gpurun_out/sanitizer.txt; : > $out
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$tool.log | tail -1)" >> $out
done
cat $out
