#!/bin/bash
# final-tree check on one GPU: all GPU tests, smoke(), the default bench line
mkdir -p gpurun_out
TAG=${1:-r2y}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_1gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'e2e',round(d['e2e']['value'],1),'pair_ms',round(r['kernel_ms'],4),'rebuild_ms',round(r['rebuild_ms_avg'],4),'C2',round(d['secondary']['C2']['steps_per_s']),'C3',d['secondary']['C3'].get('pair_kernel_ms'),d['secondary']['C3'].get('list_build_ms'),'C5',d['secondary']['C5'].get('pair_evals_per_s'))
PY
