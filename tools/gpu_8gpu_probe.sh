#!/bin/bash
# 8-GPU probes of the device-resident step: where do the ~0.09 ms per step go that one GPU does not pay?
N=${1:-8}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --no-cpu $BENCH_ARGS > gpurun_out/probe_${N}gpu_$tag.json 2> gpurun_out/probe_${N}gpu_$tag.err; echo "$tag exit $?"; \
   python -c "
import json,sys
j=json.load(open('gpurun_out/probe_${N}gpu_$tag.json'))
print('$tag', 'ms/step %.4f' % j['ms_per_step'], 'value %.0f' % j['value'], 'e2e %.0f' % j['e2e']['value'], 'e2e_base %.0f' % j['e2e_base_frames']['value'], 'parity', (j['parity'] or {}).get('ok'), 'lat', j.get('single_frame_latency_us'))
"; }
BENCH_ARGS="--steps 20 --warmup 3" run p2p_20 A=1
BENCH_ARGS="--steps 200 --warmup 3" run p2p_200 A=1
BENCH_ARGS="--steps 200 --warmup 3" run p2p_200_3bufs PISLAM_BENCH_GATHER_BUFS=3
BENCH_ARGS="--steps 200 --warmup 3 --no-parity" run noexchange_200 PISLAM_B200_DEBUG_NO_EXCHANGE=1
BENCH_ARGS="--steps 200 --warmup 3" run nccl_200 PISLAM_B200_NO_P2P=1
BENCH_ARGS="--steps 50 --warmup 3 --frames 1" run config4_frames1 A=1
