#!/bin/bash
# Decomposed path on an N-GPU box: the multi-GPU parity tests that fit, then bench.py exactly as the driver launches it
# (--steps 20 --warmup 5) and a long run, for every rank count <= N.  usage: gpu_r2_dd.sh <tag> <N> [notests]
mkdir -p gpurun_out
TAG=${1:-r2dd}; N=${2:-2}
if [ "$3" != "notests" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_dd_$TAG.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/pytest_dd_$TAG.log; tail -4 gpurun_out/pytest_dd_$TAG.log
fi
for n in 2 4 8; do
  [ $n -le $N ] || continue
  for cfg in "20 5" "300 50"; do
    set -- $cfg
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
       bench.py --gpus $n --steps $1 --warmup $2 > gpurun_out/bench_${TAG}_${n}gpu_k$1.json 2> gpurun_out/bench_${TAG}_${n}gpu_k$1.err
    echo "bench n=$n k=$1 rc=$?"; cut -c1-400 gpurun_out/bench_${TAG}_${n}gpu_k$1.json; tail -3 gpurun_out/bench_${TAG}_${n}gpu_k$1.err | cut -c1-300
  done
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/bench_${TAG}_1gpu_k20.json 2> gpurun_out/bench_${TAG}_1gpu_k20.err; echo "bench n=1 rc=$?"; cut -c1-300 gpurun_out/bench_${TAG}_1gpu_k20.json
