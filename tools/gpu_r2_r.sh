#!/bin/bash
# pair_tile.cu (TMA-staged pair kernel) with 2 / 3 / 4 tiles in flight against the gather kernel, whole-step numbers
mkdir -p gpurun_out
TAG=${1:-r2r}
for cfg in "0 0" "1 0" "1 3" "1 4"; do
  set -- $cfg
  timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu --no-e2e --no-secondary --opt pair_tile=$1 --opt pair_tile_stages=$2 > gpurun_out/bench_${TAG}_pt$1s$2.json 2> gpurun_out/bench_${TAG}_pt$1s$2.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_pt$1s$2.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('pair_tile $1 stages $2: value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4), 'list MB', d['config']['l2_policy'][:60])
PY
done
