#!/bin/bash
# BASELINE configs 3 and 5 on one GPU: 1080p and 4K pyramids, batch sweep (config 2 = the default bench line)
mkdir -p gpurun_out; : > gpurun_out/sweep_r01.jsonl
for cfg in "vga 1" "vga 16" "vga 64" "vga 256" "fhd 1" "fhd 8" "fhd 32" "uhd 1" "uhd 4" "uhd 16"; do
  set -- $cfg
  timeout 600 python bench.py --layout $1 --frames $2 --steps 20 --warmup 3 2>/dev/null >> gpurun_out/sweep_r01.jsonl
done
python - <<'PY'
import json
for ln in open("gpurun_out/sweep_r01.jsonl"):
    j = json.loads(ln)
    c = j["config"]
    print("%-4s frames %3d  kp/frame %6.0f  device %8.0f Mpix/s (%.3f ms)  e2e %7.0f Mpix/s  front frac %.3f  latency %s us" % (
        c["workload"][:3], c["frames_per_step_per_gpu"], c["keypoints_per_frame"], j["value"], j["ms_per_step"],
        j["e2e"]["value"], j["roofline"]["frac"], j["single_frame_latency_us"] and round(j["single_frame_latency_us"])))
PY
