#!/bin/bash
# One-GPU round check on the box: GPU parity tests, the bench line (both arms), the ncu launch list and
# one --set full capture of the pair kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 1000 --warmup 200 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_1gpu.json
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 60 --warmup 30 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_force_kernel -s 40 -c 2 \
    -o gpurun_out/pair_force_$TAG -f python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
if [ -n "$MC_BASELINE_ROWS_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_build_kernel -s 4 -c 1 \
    -o gpurun_out/rows_build_$TAG -f python bench.py --steps 40 --warmup 150 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_rb_$TAG.log 2>&1
echo "ncu rows_build rc=$?"
fi
ls -la gpurun_out | tail -15
