#!/bin/bash
# rows_build_kernel vs tile_build_kernel: parity on hardware, A/B bench, ncu of the new kernel
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/pytest_build_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_build_$TAG.log
for v in 1 2; do
  timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --no-e2e --no-secondary --opt build_variant=$v > gpurun_out/bench_${TAG}_v$v.json 2> gpurun_out/bench_${TAG}_v$v.err
  echo "bench v$v rc=$?"; python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_v$v.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_build_kernel -s 4 -c 1 \
    -o gpurun_out/rows_build_$TAG -f python bench.py --steps 40 --warmup 150 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_rb_$TAG.log 2>&1
echo "ncu rc=$?"
