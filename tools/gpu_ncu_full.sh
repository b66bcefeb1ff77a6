#!/bin/bash
# One ncu --set full capture of a kernel (regex) from a short bench run.  usage: gpu_ncu_full.sh <regex> <tag> [frames] [skip]
RE=${1:-front_tile_kernel}; TAG=${2:-r02}; FR=${3:-256}; SKIP=${4:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --frames $FR --no-cpu --no-parity > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
