#!/bin/bash
# One gpurun call: parity tests, smoke, a short bench (both arms) and the ncu launch list of the same command.
# usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"
tail -5 gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -3 gpurun_out/ncu_bench_$TAG.log
