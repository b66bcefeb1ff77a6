#!/bin/bash
# One ncu --set full capture of a kernel (regex) from a short bench run.  usage: gpu_ncu_full.sh <regex> <tag> [frames]
RE=${1:-front_tile_kernel}; TAG=${2:-r01}; FR=${3:-256}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --frames $FR > gpurun_out/ncu_full_$TAG.log 2>&1
tail -5 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
