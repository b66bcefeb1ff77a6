#!/bin/bash
# Round-end validation in one gpurun call: parity tests, smoke, bench (both arms), ncu launch list, a short sweep of
# the other BASELINE configs, compute-sanitizer.   usage: gpurun --timeout 1200 -- bash tools/gpu_final.sh <tag>
TAG=${1:-r02g}
bash tools/gpu_check.sh $TAG
bash tools/gpu_sweep.sh 1 ${TAG} "vga:1 vga:8 vga:64 fhd:1 fhd:8 fhd:64 uhd:1 uhd:4 uhd:16 uhd:64"
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$TAG.log | tail -1)  $(grep -c 'match True' gpurun_out/sanitizer_${tool}_$TAG.log) matches, $(grep -c 'match False' gpurun_out/sanitizer_${tool}_$TAG.log) mismatches"
done
