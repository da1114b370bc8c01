#!/bin/bash
# decomposed e2e leg with host-side phase traces (MC_E2E_TRACE: the three calls of a step; MC_TRACE_STEP: phases inside mc_step)
mkdir -p gpurun_out
TAG=${1:-r2l}; N=${2:-2}
for n in 2 4; do
  [ $n -le $N ] || continue
  MC_E2E_TRACE=1 MC_TRACE_STEP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
     bench.py --gpus $n --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_${TAG}_${n}gpu.json 2> gpurun_out/bench_${TAG}_${n}gpu.err
  echo "bench n=$n rc=$?"; grep -h "trace" gpurun_out/bench_${TAG}_${n}gpu.err | cut -c1-400
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_${n}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('n',d['n_gpus'],'value',round(d['value'],1),'steady',round(d['value_steady']['value'],1),'e2e',round(d['e2e']['value'],1), 'e2e ms',round(d['e2e']['ms_per_step'],4),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4), r['per_rank'])
PY
done
MC_E2E_TRACE=1 MC_TRACE_STEP=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
echo "bench n=1 rc=$?"; grep -h "trace" gpurun_out/bench_${TAG}_1gpu.err | cut -c1-400; cut -c1-200 gpurun_out/bench_${TAG}_1gpu.json
