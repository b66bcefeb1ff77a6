#!/bin/bash
# Batch-size sweep of the other BASELINE configs (config 3 = 1080p, config 4 = one VGA pyramid per
# GPU + gather, config 5 = 4K, batch 1..256) on N GPUs of one box; one bench.py JSON line per run (parity gate on).
# usage: gpurun [--gpus N] --timeout 3000 -- bash tools/gpu_sweep.sh N tag "layout:frames ..."
N=${1:-1}; TAG=${2:-r02}; LIST=${3:-"vga:1 vga:8 vga:64 vga:256 fhd:1 fhd:8 fhd:64 uhd:1 uhd:2 uhd:4 uhd:8 uhd:16 uhd:32 uhd:64 uhd:128 uhd:256"}
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${TAG}_${N}gpu.jsonl
: > $OUT
for item in $LIST; do
  L=${item%%:*}; F=${item##*:}
  STEPS=20; [ "$L" = "uhd" ] && [ "$F" -ge 64 ] && STEPS=5
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --layout $L --frames $F --steps $STEPS --warmup 3 --no-cpu >> $OUT 2>> gpurun_out/sweep_${TAG}_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --layout $L --frames $F --steps $STEPS --warmup 3 --no-cpu >> $OUT 2>> gpurun_out/sweep_${TAG}_${N}gpu.err
  fi
  echo "$item exit $?"
done
python - <<PY
import json
for ln in open("$OUT"):
    try: j = json.loads(ln)
    except Exception: continue
    print("%-4s F=%-4d N=%d  device %9.1f Gpix/s  e2e %7.1f  e2e_base %7.1f  ms/step %8.3f  front frac %.3f  parity %s  kp/frame %.0f" % (
        j["config"]["workload"][:3], j["config"]["frames_per_step_per_gpu"], j["n_gpus"], j["value"]/1e3, j["e2e"]["value"]/1e3,
        j["e2e_base_frames"]["value"]/1e3, j["ms_per_step"], (j["roofline"] or {}).get("frac", 0), (j["parity"] or {}).get("ok"), j["keypoints_per_frame"]))
PY
