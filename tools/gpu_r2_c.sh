#!/bin/bash
# pair_tile.cu iteration on one B200: parity suite, bench (tile kernel, with the C2/C3/C5 blocks), gather kernel for
# reference, ncu --set full of the tile kernel.  usage: gpu_r2_c.sh <tag> [extra bench flags]
mkdir -p gpurun_out
TAG=${1:-r2c}; shift
B="--steps 300 --warmup 50 --no-cpu --no-e2e"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py $B "$@" > gpurun_out/bench_${TAG}_tile.json 2> gpurun_out/bench_${TAG}_tile.err; echo "tile rc=$?"; cat gpurun_out/bench_${TAG}_tile.json; tail -2 gpurun_out/bench_${TAG}_tile.err
timeout 600 python bench.py $B --no-secondary --opt pair_tile=0 > gpurun_out/bench_${TAG}_gather.json 2> gpurun_out/bench_${TAG}_gather.err; echo "gather rc=$?"; cat gpurun_out/bench_${TAG}_gather.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 30 -c 1 \
    -o gpurun_out/pair_tile_$TAG -f python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
