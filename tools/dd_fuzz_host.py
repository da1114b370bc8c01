#!/usr/bin/env python
"""Randomised sweep of the decomposed path on the HOST build of the library (no GPU): rank counts 2-4, system sizes,
temperatures up to 900 K, skins, fixed / adaptive / all-gather schedules, NCCL and fused peer-memory halo; every run is
compared with the oracle's trajectory.  Refusals ("need >= 2 layers per rank") are printed as FAIL with the message.
    bash tests/cpp/host_lib/build.sh
    MOLCHANICA_MD_LIB=$PWD/tests/cpp/_build/libmolchanica_md_host.so MOLCHANICA_NCCL_LIB=$PWD/tests/cpp/_build/libnccl_standin.so \
        MC_SHIM_THREADS=2 python tools/dd_fuzz_host.py <seed> <trials>
Round 1: 36 configurations, all valid ones equal to the oracle, no list violations."""
import os, sys, time, subprocess, tempfile, itertools
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import trajectory_close
from oracle import oracle_py as oracle
def run(world, halo, sched, env):
    d = tempfile.mkdtemp()
    idf, out = os.path.join(d, "nccl_id"), os.path.join(d, "out.npz")
    e = dict(os.environ, **env, MC_SHIM_SHARED_HEAP="1" if halo == "fused" else "0")
    procs = [subprocess.Popen([sys.executable, "tests/dd_worker.py", str(r), str(world), idf, "ljx", out, halo, sched],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e) for r in range(world)]
    logs = [p.communicate(timeout=900)[0] for p in procs]
    if not all(p.returncode == 0 for p in procs): return None, "\n".join(l[-400:] for l in logs)
    return np.load(out), ""
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    world = int(rng.choice([2, 3, 4]))
    m = int(rng.choice([20, 24, 28]))
    temp = float(rng.choice([86.3, 300.0, 900.0]))
    skin = float(rng.choice([0.6, 1.0, 1.6]))
    every = int(rng.choice([2, 4, 7]))
    halo = str(rng.choice(["fused", "nccl"]))
    sched = str(rng.choice(["fixed", "adaptive", "allgather"]))
    steps = int(rng.choice([25, 50]))
    env = dict(DD_M=str(m), DD_TEMP=str(temp), DD_SKIN=str(skin), DD_STEPS=str(steps), DD_EVERY=str(every))
    os.environ.update(env)
    from dd_worker import case_workload
    w, n_steps = case_workload("ljx", world)
    t = time.time()
    r, err = run(world, halo, sched, env)
    if r is None:
        print("FAIL", world, m, temp, skin, every, halo, sched, steps, err[-600:]); continue
    ref = oracle.md_run(w, n_steps, precision=64)
    ok, worst, sc = trajectory_close(r["x"], ref["xyzq"], w["xyzq"], w["box_ext"])
    print("ok " if ok and int(r["violations"]) == 0 and bool(r["snap_ok"]) else "BAD", "world", world, "m", m, "T", temp, "skin", skin, "every", every, halo, sched, "steps", steps,
          "worst %.2e" % worst, "viol", int(r["violations"]), "interval", int(r["interval"]), "rebuilds", int(r["rebuilds"]), "disp %.2f" % float(r["disp_frac"]), "%.0fs" % (time.time() - t))
