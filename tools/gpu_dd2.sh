#!/bin/bash
# Two-GPU check of the decomposed path: parity tests (fused + NCCL halo), then the bench at N=2 both ways.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_dd2_$TAG.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_dd2_$TAG.log
for mode in 1 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 1000 --warmup 200 --opt halo_fused=$mode > gpurun_out/bench_${TAG}_2gpu_fused$mode.json 2> gpurun_out/bench_${TAG}_2gpu_fused$mode.err
  echo "bench fused=$mode rc=$?"; cat gpurun_out/bench_${TAG}_2gpu_fused$mode.json; tail -5 gpurun_out/bench_${TAG}_2gpu_fused$mode.err
done
