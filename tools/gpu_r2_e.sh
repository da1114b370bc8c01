#!/bin/bash
# bench only (no tests): default build + variant libraries.  usage: gpu_r2_e.sh <tag>
mkdir -p gpurun_out
TAG=${1:-r2x}
B="--steps 200 --warmup 50 --no-cpu --no-e2e --no-secondary --no-steady"
timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_tile.json 2> gpurun_out/bench_${TAG}_tile.err; echo "tile rc=$?"; cat gpurun_out/bench_${TAG}_tile.json | cut -c1-300; tail -2 gpurun_out/bench_${TAG}_tile.err
for V in molchanica_b200/_variants/libmolchanica_md_*.so; do
  [ -f $V ] || continue
  v=$(basename $V .so | sed 's/libmolchanica_md_//')
  MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 MOLCHANICA_MD_LIB=$PWD/$V timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  echo "$v rc=$?"; cat gpurun_out/bench_${TAG}_$v.json | cut -c1-300
done
