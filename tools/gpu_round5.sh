#!/bin/bash
# ncu --set full of the tile kernel (shipped build) + A/B of variants
mkdir -p gpurun_out
bash tools/gpu_ncu_full.sh front_tile_kernel r02_front_v5 256 3
timeout 600 python tools/gpu_front_ab.py ab/lib_base.so ab/lib_new.so ab/lib_c8.so > gpurun_out/ab_r02_v5.txt 2>&1
cat gpurun_out/ab_r02_v5.txt
