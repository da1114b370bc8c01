#!/bin/bash
# 2-GPU box: stall hunt in the end-to-end leg (MC_TRACE_STEP + MC_TRACE_ALLOC: allocations and phases / rebuilds over 5 ms on stderr)
mkdir -p gpurun_out
TAG=${1:-r2w}
for k in 300 300; do
MC_TRACE_STEP=1 MC_TRACE_ALLOC=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps $k --warmup 20 --no-steady \
     > gpurun_out/bench_${TAG}_2gpu_k$k.json 2> gpurun_out/bench_${TAG}_2gpu_k$k.err
grep -n "stall\|mc alloc" gpurun_out/bench_${TAG}_2gpu_k$k.err | tail -40
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_2gpu_k$k.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print('2 GPUs K=$k: value',round(d['value'],1),'e2e',round(e['value'],1),'worst',e['worst_step_ms'],e['worst_step_index'])
PY
done
