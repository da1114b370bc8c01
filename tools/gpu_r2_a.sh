#!/bin/bash
# Round 2, first call (one B200): GPU parity suite, the bench line under the driver's own flags, the reference arm,
# the long default bench line, and the A/B of the tile-build register budget (variant library built beforehand by
# tools/gpu_ab_variant.sh's make line; the .so travels).  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_1gpu_k20.json 2> gpurun_out/bench_${TAG}_1gpu_k20.err
echo "bench k20 rc=$?"; cat gpurun_out/bench_${TAG}_1gpu_k20.json; tail -3 gpurun_out/bench_${TAG}_1gpu_k20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_ref.json
timeout 600 python bench.py --no-cpu --no-secondary > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_1gpu.json; tail -3 gpurun_out/bench_${TAG}_1gpu.err
V=molchanica_b200/_variants/libmolchanica_md_tile2.so
if [ -f $V ]; then
  MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 MOLCHANICA_MD_LIB=$PWD/$V timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --no-secondary --no-e2e > gpurun_out/bench_${TAG}_tile2.json 2> gpurun_out/bench_${TAG}_tile2.err
  echo "tile2 rc=$?"; cat gpurun_out/bench_${TAG}_tile2.json
fi
