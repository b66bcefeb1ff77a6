#!/bin/bash
# scaling run on an N-GPU box: bench at 1,2,4,..,N GPUs (our arm) + the reference arm once
N=${1:-8}; TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_gpus_$TAG.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/scale_ref_$TAG.json 2>/dev/null
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_1_$TAG.json 2> gpurun_out/scale_1_$TAG.err
for g in 2 4 8; do
  if [ $g -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g \
      bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/scale_${g}_$TAG.json 2> gpurun_out/scale_${g}_$TAG.err
    echo "exit $?" >> gpurun_out/scale_${g}_$TAG.err
  fi
done
for g in 1 2 4 8; do [ -f gpurun_out/scale_${g}_$TAG.json ] && python - <<PY
import json
try:
    j=json.load(open("gpurun_out/scale_${g}_$TAG.json"))
    print(j["n_gpus"], "GPUs: value %.0f Mpix/s, %.3f ms/step, e2e %.0f Mpix/s, %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["config"]["gather"]))
except Exception as e:
    print("$g: no result", e)
PY
done
tail -3 gpurun_out/scale_8_$TAG.err 2>/dev/null
