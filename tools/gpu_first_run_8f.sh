#!/bin/bash
# First hardware run of the SURVEY 8f components that were written after round 1's GPU budget was spent
# (bonded.cu, settle.cu incl. virtual sites, thermostat.cu, pme.cu).  One GPU.  Each worker prints one JSON line
# with the measured errors; nothing here may hang: every step has its own timeout.
mkdir -p gpurun_out
TAG=${1:-8f}
# everything added after the last hardware run in one go first: the parity file (pipelined external forces, pressure,
# constraint virial, barostat, snapshots with velocities, drift removal) -- these already pass against the host build
timeout 1200 python -m pytest tests/test_gpu_md_paths.py tests/test_gpu_edge_cases.py -m gpu -q > gpurun_out/new_paths_$TAG.log 2>&1; echo "new paths rc=$?"; tail -5 gpurun_out/new_paths_$TAG.log
# the e2e line with and without the pipelined upload (option defer_tail): the A/B that says what it bought
timeout 600 python bench.py --steps 300 --warmup 50 > gpurun_out/bench_defer_on_$TAG.json 2> gpurun_out/bench_defer_on_$TAG.err; tail -1 gpurun_out/bench_defer_on_$TAG.json
timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --opt defer_tail=0 > gpurun_out/bench_defer_off_$TAG.json 2> gpurun_out/bench_defer_off_$TAG.err; tail -1 gpurun_out/bench_defer_off_$TAG.json
for w in bonded settle langevin pme; do
  echo "== $w"
  timeout 600 python tests/${w}_gpu_worker.py > gpurun_out/${w}_worker_$TAG.json 2> gpurun_out/${w}_worker_$TAG.err
  echo "rc=$?"; cat gpurun_out/${w}_worker_$TAG.json; tail -5 gpurun_out/${w}_worker_$TAG.err
done
timeout 300 compute-sanitizer --tool memcheck python tests/bonded_gpu_worker.py > gpurun_out/sanitizer_bonded_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_bonded_$TAG.log
timeout 300 compute-sanitizer --tool memcheck python tests/pme_gpu_worker.py > gpurun_out/sanitizer_pme_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_pme_$TAG.log
timeout 300 compute-sanitizer --tool racecheck python tests/settle_gpu_worker.py > gpurun_out/racecheck_settle_$TAG.log 2>&1; tail -3 gpurun_out/racecheck_settle_$TAG.log
timeout 900 python tools/measure_configs.py > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err; cat gpurun_out/configs_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_8f_$TAG.csv \
    python tests/pme_gpu_worker.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_8f_$TAG.csv | head -20
