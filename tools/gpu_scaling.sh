#!/bin/bash
# On an 8-GPU box: decomposed parity tests, then the bench at N = 1, 2, 4, 8 (fused halo) and N = 8 with NCCL halo.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_dd_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dd_$TAG.log
timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
cat gpurun_out/bench_${TAG}_1gpu.json
for cfg in "2 1" "4 1" "8 1" "8 0"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 \
     bench.py --gpus $1 --steps 1000 --warmup 200 --opt halo_fused=$2 > gpurun_out/bench_${TAG}_$1gpu_fused$2.json 2> gpurun_out/bench_${TAG}_$1gpu_fused$2.err
  echo "bench N=$1 fused=$2 rc=$?"; cat gpurun_out/bench_${TAG}_$1gpu_fused$2.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_${TAG}_$1gpu_fused$2.err | tail -5
done
