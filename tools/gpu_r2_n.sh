#!/bin/bash
# all GPU tests (no -x: every failure at once), then phase time stamps of the fused kernel on C2, wall time per call
mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$TAG.log
MC_FUSED_TIMES=1 timeout 300 python - > gpurun_out/fused_times_$TAG.log 2>&1 <<'PY'
import sys, time; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.globule(); e=MdEngine.from_workload(w)
for k in range(6): e.step(w["dt"],10)
PY
tail -3 gpurun_out/fused_times_$TAG.log | cut -c1-900
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.globule(); e=MdEngine.from_workload(w)
for k in range(50): e.step(w["dt"],10)
t0=time.perf_counter()
for k in range(500): e.step(w["dt"],10)
t=time.perf_counter()-t0
print("C2 wall us/call", t/500*1e6, "last_step_ms", e.last_step_ms())
t0=time.perf_counter()
for k in range(50): e.step(w["dt"],100)
t=time.perf_counter()-t0
print("C2 100-step calls: us/step", t/5000*1e6, "last_step_ms", e.last_step_ms())
PY
