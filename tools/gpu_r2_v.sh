#!/bin/bash
# 2-GPU box: decomposed parity (all world-2 cases incl. bonded / Langevin / pressure on decomposed handles), then the bench
# exactly as the driver launches it (K = 20, W = 5) and with a long window
mkdir -p gpurun_out
TAG=${1:-r2v}
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_dd_$TAG.log 2>&1; echo "dd tests rc=$?"; tail -4 gpurun_out/pytest_dd_$TAG.log
for cfg in "20 5" "300 50"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps $1 --warmup $2 \
     > gpurun_out/bench_${TAG}_2gpu_k$1.json 2> gpurun_out/bench_${TAG}_2gpu_k$1.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_2gpu_k$1.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('2 GPUs K=$1: value',round(d['value'],1),'steady',round(d['value_steady']['value'],1) if d.get('value_steady') else None,'e2e',round(d['e2e']['value'],1),'pair_ms',r['per_rank']['pair_ms'],'integrate',r['per_rank']['integrate_ms'],'rebuild',r['per_rank']['rebuild_ms'], 'interval', r['per_rank']['rebuild_interval'])
PY
done
