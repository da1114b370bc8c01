#!/bin/bash
# quick 1-GPU check: GPU tests + bench line (+ optional microbench)
mkdir -p gpurun_out
TAG=${1:-chk}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 1000 --warmup 200 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_1gpu.json; tail -3 gpurun_out/bench_${TAG}_1gpu.err
[ -x tools/microbench/ffma2 ] && ./tools/microbench/ffma2 | tee gpurun_out/ffma2_$TAG.txt
