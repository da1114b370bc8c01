#!/bin/bash
# fused multi-step kernel first (bounded: its first hardware run hung), then the whole GPU suite and a bench line
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 300 python -m pytest tests/test_gpu_md_paths.py -m gpu -x -q -k fused > gpurun_out/pytest_fused_$TAG.log 2>&1
rc=$?; echo "fused rc=$rc"; tail -5 gpurun_out/pytest_fused_$TAG.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
