for pad in 0 3000 8000 15000; do PISLAM_B200_DEBUG_SMEM_PAD=$pad python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('pad $pad', 'front %.4f' % j['roofline']['avg_launch_ms'])"; done
