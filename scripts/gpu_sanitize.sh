#!/bin/bash
mkdir -p gpurun_out/r02c
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 8 python scripts/sanitizer_workload.py > gpurun_out/r02c/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|workload done" gpurun_out/r02c/sanitizer_$tool.log | tail -3
  grep -E "Race reported|Invalid|hazard|Barrier error|divergent" gpurun_out/r02c/sanitizer_$tool.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -8
done
