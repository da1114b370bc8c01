#!/bin/bash
# 2-GPU box: whole GPU suite (incl. the decomposed tests that fit 2 GPUs), then the decomposed e2e leg with host-side traces
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_md_paths.py -m gpu -x -q > gpurun_out/pytest_dd_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dd_$TAG.log
true
bash tools/gpu_r2_l.sh $TAG 2
true <<'PY'
import sys; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.globule(); e=MdEngine.from_workload(w)
for k in range(30): e.step(w["dt"],10)
PY
echo "ncu fused rc=$?"; tail -8 gpurun_out/ncu_fused_$TAG.csv | cut -c1-300
