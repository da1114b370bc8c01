#!/bin/bash
# pair_tile.cu iteration (one B200): quick parity subset, bench of the default build and of variant libraries, one ncu capture.
mkdir -p gpurun_out
TAG=${1:-r2d}; shift
B="--steps 300 --warmup 50 --no-cpu --no-e2e --no-secondary"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_tile.json 2> gpurun_out/bench_${TAG}_tile.err; echo "tile rc=$?"; cat gpurun_out/bench_${TAG}_tile.json; tail -2 gpurun_out/bench_${TAG}_tile.err
for V in molchanica_b200/_variants/libmolchanica_md_*.so; do
  [ -f $V ] || continue
  v=$(basename $V .so | sed 's/libmolchanica_md_//')
  MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 MOLCHANICA_MD_LIB=$PWD/$V timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  echo "$v rc=$?"; cat gpurun_out/bench_${TAG}_$v.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 30 -c 1 \
    -o gpurun_out/pair_tile_$TAG -f python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
