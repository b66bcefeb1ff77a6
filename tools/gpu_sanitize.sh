#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)  $(grep -c 'match True' gpurun_out/sanitizer_$tool.log) matches, $(grep -c 'match False' gpurun_out/sanitizer_$tool.log) mismatches"
done
