#!/bin/bash
# compile-time knobs of pair_force.cu's row loop (built here: molchanica_b200/_variants/libmolchanica_md_<tag>.so) against the default
mkdir -p gpurun_out
TAG=${1:-r2t}; shift
for v in base "$@"; do
  if [ $v = base ]; then unset MOLCHANICA_MD_LIB MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE; else export MOLCHANICA_MD_LIB=$PWD/molchanica_b200/_variants/libmolchanica_md_$v.so MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1; fi
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact" > gpurun_out/pytest_${TAG}_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/pytest_${TAG}_$v.log)"
  timeout 300 python bench.py --steps 200 --warmup 50 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_$v.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$v: value',round(d['value'],1),'rebuild_ms',round(r['rebuild_ms_avg'],4),'pair_ms',round(r['kernel_ms'],4))
PY
done
