#!/bin/bash
# e2e leg under the three ways a pipelined call can end: lazy return + early force launch, synchronised + early launch, neither
mkdir -p gpurun_out
TAG=${1:-r2p}; N=${2:-2}
for n in 1 $N; do
 for cfg in "1 1" "0 1" "0 0"; do
  set -- $cfg
  if [ $n -gt 1 ]; then
    MC_E2E_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n \
       bench.py --gpus $n --steps 200 --warmup 20 --no-cpu --no-secondary --no-steady --opt lazy_sync=$1 --opt early_tail=$2 > gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.json 2> gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.err
  else
    MC_E2E_TRACE=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary --no-steady --opt lazy_sync=$1 --opt early_tail=$2 > gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.json 2> gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.err
  fi
  grep -h "e2e trace" gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.err | cut -c1-200
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_${n}gpu_l$1e$2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('n',d['n_gpus'],'lazy $1 early $2: value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['e2e']['ms_per_step'],4),'rebuilds',d['e2e']['rebuilds'])
PY
 done
done
