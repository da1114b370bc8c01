#!/bin/bash
# Round 2, pair_tile.cu (TMA-staged force kernel) on one B200: parity suite, bench A/B against the gather kernel and two
# CTA shapes, ncu --set full of the new kernel, launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2b}
B="--steps 300 --warmup 50 --no-cpu --no-secondary --no-e2e"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_tile.json 2> gpurun_out/bench_${TAG}_tile.err; echo "tile rc=$?"; cat gpurun_out/bench_${TAG}_tile.json; tail -2 gpurun_out/bench_${TAG}_tile.err
timeout 600 python bench.py $B --opt pair_tile=0 > gpurun_out/bench_${TAG}_gather.json 2> gpurun_out/bench_${TAG}_gather.err; echo "gather rc=$?"; cat gpurun_out/bench_${TAG}_gather.json
for v in pt4 pt8b4; do
  V=$PWD/molchanica_b200/_variants/libmolchanica_md_$v.so
  [ -f $V ] || continue
  MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 MOLCHANICA_MD_LIB=$V timeout 600 python bench.py $B > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  echo "$v rc=$?"; cat gpurun_out/bench_${TAG}_$v.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel -s 30 -c 2 \
    -o gpurun_out/pair_tile_$TAG -f python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 60 --warmup 30 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | tail -12
