#!/bin/bash
# quad-interleaved rows (option rows_interleave) against plain rows: bit-identity test, then whole-step numbers
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "interleaved or bit_exact" > gpurun_out/pytest_ilv_$TAG.log 2>&1
tail -3 gpurun_out/pytest_ilv_$TAG.log
for ilv in 1 0; do
  timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --no-e2e --no-secondary --opt rows_interleave=$ilv > gpurun_out/bench_${TAG}_ilv$ilv.json 2> gpurun_out/bench_${TAG}_ilv$ilv.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_ilv$ilv.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('rows_interleave $ilv: value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4), 'integrate_ms', round(r['integrate_ms_avg'],4))
PY
  tail -2 gpurun_out/bench_${TAG}_ilv$ilv.err
done
