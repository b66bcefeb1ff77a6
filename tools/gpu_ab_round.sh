#!/bin/bash
# One gpurun call: A/B of prebuilt library variants (ab/lib_*.so), then the GPU parity tests against one of them.
# usage: gpurun --timeout 900 -- bash tools/gpu_ab_round.sh <tag> <lib for pytest or -> <lib> <lib> ...
TAG=$1; PYLIB=$2; shift 2
mkdir -p gpurun_out
timeout 600 python tools/gpu_front_ab.py "$@" > gpurun_out/ab_$TAG.txt 2>&1
cat gpurun_out/ab_$TAG.txt
if [ "$PYLIB" != "-" ]; then
  PISLAM_B200_LIB=$PWD/$PYLIB timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest ($PYLIB) exit $?" >> gpurun_out/pytest_$TAG.log
  tail -15 gpurun_out/pytest_$TAG.log
fi
