#!/bin/bash
# fused kernel (brute-force private list) bounded first, whole GPU suite, bench line with the secondary configs, ncu of rows_build_kernel
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 300 python -m pytest tests/test_gpu_md_paths.py -m gpu -x -q -k fused > gpurun_out/pytest_fused_$TAG.log 2>&1
rc=$?; echo "fused rc=$rc"; tail -5 gpurun_out/pytest_fused_$TAG.log
if [ $rc -eq 0 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
  timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
  python - <<PY
import json
for l in open('gpurun_out/bench_$TAG.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'e2e',round(d['e2e']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4))
        print(json.dumps(d['secondary'])[:900])
PY
fi
for mb in 2 3; do
  timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --no-e2e --no-secondary --opt rows_min_blocks=$mb > gpurun_out/bench_${TAG}_mb$mb.json 2> gpurun_out/bench_${TAG}_mb$mb.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_mb$mb.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('min_blocks $mb: value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4))
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_build_kernel -s 4 -c 1 \
    -o gpurun_out/rows_build_$TAG -f python bench.py --steps 40 --warmup 150 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_rb_$TAG.log 2>&1
echo "ncu rc=$?"
