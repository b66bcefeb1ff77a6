for c in 2 4 8; do PISLAM_B200_CHUNKS=$c python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('chunks $c e2e %.3f ms  %.0f Mpix/s' % (j['e2e']['ms_per_step'], j['e2e']['value']))"; done
PISLAM_B200_TRIM=0 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('notrim e2e %.3f ms  %.0f Mpix/s' % (j['e2e']['ms_per_step'], j['e2e']['value']))"
