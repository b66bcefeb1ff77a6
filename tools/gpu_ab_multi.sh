#!/bin/bash
# bench every ab/lib_*.so once (device-resident numbers only)
for f in ab/lib_*.so; do
PISLAM_B200_LIB=$PWD/$f python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$f', 'step %.4f ms' % j['ms_per_step'], 'front %.4f' % j['roofline']['avg_launch_ms'], 'orb %.4f' % j['roofline']['other_kernels_avg_ms']['orb_kernel'], 'e2e %.3f' % j['e2e']['ms_per_step'])"
done
