#!/bin/bash
# fused kernel: tests of the file that covers it, then wall time per call on C2 (10- and 100-step calls), lanes sweep
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.globule(temp_k=100.0); w["dt"]=0.0002   # bench.py's C2: no bonded terms, a short step keeps the globule intact
for opts in ({}, {"fused_lanes":8}, {"fused_lanes":16}, {"fused_brute":0}, {"fused_steps":0}):
    e=MdEngine.from_workload(w)
    for k,v in opts.items(): e.set_option(k,v)
    for k in range(100): e.step(w["dt"],10)
    t0=time.perf_counter()
    for k in range(500): e.step(w["dt"],10)
    t=time.perf_counter()-t0
    t1=time.perf_counter()
    for k in range(50): e.step(w["dt"],100)
    t2=time.perf_counter()-t1
    print(opts, "10-step calls: us/step %.2f (%.0f steps/s)" % (t/5000*1e6, 5000/t), " 100-step calls: us/step %.2f" % (t2/5000*1e6), "rebuilds", e.stats()["n_rebuilds"])
    e.close()
PY
MC_FUSED_TIMES=1 timeout 300 python - 2>&1 <<'PY' | tail -2 | cut -c1-700
import sys, time; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.globule(temp_k=100.0); w["dt"]=0.0002; e=MdEngine.from_workload(w)
import os
os.environ.pop("MC_FUSED_TIMES",None)
for k in range(300): e.step(w["dt"],10)
PY
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
w=W.bonded_globule(1231, seed=202)
for opts in ({}, {"fused_steps":0}):
    e=MdEngine.from_workload(w, bonded=True)
    for k,v in opts.items(): e.set_option(k,v)
    for k in range(100): e.step(w["dt"],10)
    t0=time.perf_counter()
    for k in range(500): e.step(w["dt"],10)
    t=time.perf_counter()-t0
    print("bonded globule 2 fs", opts, "10-step calls: us/step %.2f (%.0f steps/s)" % (t/5000*1e6, 5000/t), "rebuilds", e.stats()["n_rebuilds"], "T", e.energy()["temperature"])
    e.close()
PY
