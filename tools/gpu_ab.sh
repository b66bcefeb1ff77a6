#!/bin/bash
# A/B two builds of the library in one box: ab/lib_a.so vs ab/lib_b.so (each run twice, interleaved)
for r in 1 2; do for v in a b; do
PISLAM_B200_LIB=$PWD/ab/lib_$v.so python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$v', 'step %.4f ms' % j['ms_per_step'], 'front %.4f' % j['roofline']['avg_launch_ms'], j['roofline']['other_kernels_avg_ms'], 'e2e %.3f' % j['e2e']['ms_per_step'])"
done; done
