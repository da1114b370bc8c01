#!/bin/bash
# N-GPU box: the decomposed parity tests that fit, then bench.py at N ranks exactly as the driver launches it (--steps 20 --warmup 5)
# and a long run.  usage: gpu_r2_q.sh <tag> <N> [notests]
mkdir -p gpurun_out
TAG=${1:-r2q}; N=${2:-4}
if [ "$3" != "notests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_dd_$TAG.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/pytest_dd_$TAG.log; tail -4 gpurun_out/pytest_dd_$TAG.log
fi
for cfg in "20 5" "300 50"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2981$N \
     bench.py --gpus $N --steps $1 --warmup $2 --no-cpu > gpurun_out/bench_${TAG}_${N}gpu_k$1.json 2> gpurun_out/bench_${TAG}_${N}gpu_k$1.err
  echo "bench n=$N k=$1 rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_${N}gpu_k$1.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('n',d['n_gpus'],'k',d['steps'],'value',round(d['value'],1),'steady',d['value_steady'] and round(d['value_steady']['value'],1),'e2e',round(d['e2e']['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4), r['per_rank'])
PY
done
