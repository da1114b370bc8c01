#!/bin/bash
# one B200: full GPU parity suite + default bench line (with secondary configs)
mkdir -p gpurun_out
TAG=${1:-r2x}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
