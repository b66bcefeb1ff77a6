#!/bin/bash
# Builds the library of a git revision into ab/<name>.so (for tools/gpu_front_ab.py A/B runs against the working tree).
# usage: tools/build_ref_lib.sh <git-ref> <name> [extra nvcc flags...]
set -e
REF=$1; NAME=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REF" pislam_b200/csrc include | tar -x -C "$TMP"
mkdir -p "$ROOT/ab"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o "$ROOT/ab/$NAME.so" "$TMP/pislam_b200/csrc/pislam_b200.cu" -ldl
rm -rf "$TMP"
echo "$ROOT/ab/$NAME.so"
