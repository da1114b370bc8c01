#!/usr/bin/env python
"""bench.py -- ns/day of the MD hot path on BASELINE.json's headline configuration (C4: the
1,000,000-atom LJ fluid, strong scaling over 1/2/4/8 B200) plus the pair-force kernel's
HBM-roofline fraction, next to the CPU restatement timed on the host cores.

    python bench.py --gpus N --steps K --warmup W          # this engine
    python bench.py --impl reference --steps K --warmup W  # CPU arm (oracle port)

A "step" is one velocity-Verlet MD step of the whole system: kick+drift, (Verlet-list rebuild
when the displacement criterion fires), pair forces, kick.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ns_per_day_1M_atom_lj_fluid"
UNIT = "ns/day"
DT_PS = 0.002


def ns_per_day(steps, seconds, dt_ps=DT_PS):
    return steps * dt_ps * 1e-3 / seconds * 86400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  The sampler process is
    started well before the first timed window (its start-up -- fork, exec, NVML initialisation -- would otherwise steal
    the launching thread's time inside a 20-step window of a few milliseconds) and its rows carry timestamps; only the
    rows that fall between mark_begin() and mark_end() are reported."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in ln.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        note = None
        lo, hi = self.t_begin or 0.0, (self.t_end or time.time()) + 0.03
        rows = [r for (t, r) in self.rows if lo <= t <= hi]
        if not rows and self.rows:
            # the timed regions were shorter than one sampling interval: the nearest rows around them
            rows = [r for (t, r) in sorted(self.rows, key=lambda tr: min(abs(tr[0] - lo), abs(tr[0] - hi)))[:3]]
            note = "timed regions shorter than the sampling interval; nearest samples reported"
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[2]))
                mx = float(r[3])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
               "window": "value + steady-state + end-to-end legs"}
        if note:
            out["note"] = note
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(side, steps, warmup):
    """The CPU path (oracle port, OpenMP over all host cores) on the SAME workload: side = 100 is the full
    1,000,000-atom C4 fluid; the sample is bounded in STEPS, not in atoms.  A smaller side (--cpu-side) runs the same
    fluid at the same density in a side^3-atom box and scales by atoms (the work per atom is identical); the line says
    which.  torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly to the cores this process may use."""
    from molchanica_b200 import workloads as W
    from oracle import oracle_py as O
    O.lib()
    O.set_num_threads(host_cores())
    w = W.lj_fluid(m=side)
    n = len(w["xyzq"])
    st = O.md_run(w, warmup, precision=32)
    t0 = time.perf_counter()
    O.md_run(w, steps, precision=32, xyzq=st["xyzq"], vel=st["vel"])
    dt = time.perf_counter() - t0
    v_sample = ns_per_day(steps, dt)
    v_1m = v_sample * n / 1.0e6
    full = n == 1000000
    return dict(value=v_1m, unit=UNIT, cores=O.num_threads(), kind="port", same_config=full,
                sample=(f"the full {n}-atom C4 fluid, {steps} steps incl. list rebuilds after {warmup} warm-up steps, {dt:.2f} s"
                        if full else
                        f"{n}-atom box of the same LJ fluid (side {side}), {steps} steps incl. list rebuilds, {dt:.2f} s; "
                        f"scaled by atoms to 1M"),
                seconds=dt, steps=steps)


def secondary_configs(peak_gbs):
    """BASELINE.json configs 2, 3 and 5 on this GPU (one small block each, a few seconds in total): C2 steps/s in
    batches of 10 like the GUI (reference src/md/mod.rs:45), C3 pair-force and list-build kernels individually,
    C5 pose-energy scan in pair evaluations per second against the FP32 issue ceiling."""
    from molchanica_b200 import workloads as W
    from molchanica_b200.engine import MdEngine
    out = {}
    try:
        w = W.globule(temp_k=100.0)
        e = MdEngine.from_workload(w)
        dt_run = 0.0002  # the synthetic globule has no bonded terms: a short step keeps it intact; cost per step is dt-independent
        e.step(dt_run, 200)
        s0 = e.stats()
        t0 = time.perf_counter()
        for _ in range(1000):
            e.step(dt_run, 10)
        el = time.perf_counter() - t0
        s1 = e.stats()
        out["C2"] = {"workload": f"{len(w['xyzq'])}-atom globule in vacuum, 10,000 steps in mc_step batches of 10 (one cooperative launch per "
                                 "batch: md_fused.cu, private all-pairs list rebuilt inside the launch)", "steps_per_s": 10000 / el,
                     "us_per_step": el / 10000 * 1e6, "ns_per_day_at_2fs": ns_per_day(10000, el),
                     "launches_per_step": (s1["n_kernel_launches"] - s0["n_kernel_launches"]) / 10000.0,
                     "rebuilds": int(s1["n_rebuilds"] - s0["n_rebuilds"])}
        e.close()
    except Exception as ex:  # noqa: BLE001
        out["C2"] = {"error": str(ex)}
    try:
        w = W.solvated_c3()
        e = MdEngine.from_workload(w)
        e.set_option("profiling", 1)
        for _ in range(3):
            e.build_neighbors()
        e.reset_timers()
        for _ in range(10):
            e.build_neighbors()
        sb = e.stats()
        pair_ms = e.time_pair_kernel(reps=50, flush_l2=True)
        n, p = len(w["xyzq"]), sb["n_pairs_listed"]
        alg = 32.0 * n + 20.0 * p
        build_ms = sb["build_ms_sum"] / 10.0  # ten builds; each one is bracketed in two parts (sort + reorder, rows)
        out["C3"] = {"workload": f"{n} atoms solvated, rc 12 A + 2 A skin, {p} list entries", "pair_kernel_ms": pair_ms,
                     "pair_algorithmic_GBs": alg / pair_ms / 1e6, "pair_frac_of_measured_hbm": alg / pair_ms / 1e6 / peak_gbs,
                     "list_build_ms": build_ms, "list_build_algorithmic_GBs": (120.0 * n + 4.0 * p) / max(build_ms, 1e-9) / 1e6,
                     "builds_timed": int(sb["builds_timed"])}
        e.close()
    except Exception as ex:  # noqa: BLE001
        out["C3"] = {"error": str(ex)}
    try:
        d = W.docking_c5()
        e = MdEngine()
        e.set_option("profiling", 1)
        e.dock_score(d)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            e.dock_score(d)
        wall = (time.perf_counter() - t0) / reps
        kms = e.last_dock_kernel_ms()
        pairs = len(d["poses"]) * len(d["rec"]) * len(d["lig"])
        # ceiling: 128 FP32 lanes/clk/SM x 148 SMs x max SM clock / issue slots per pair (counted in the SASS of the
        # inner loop, tools/sass_count.sh -> profiles/dock_sass_r2.txt)
        slots = DOCK_ISSUE_SLOTS_PER_PAIR
        ceil = 128.0 * 148 * 1.965e9 / slots
        out["C5"] = {"workload": f"{len(d['poses'])} poses x {len(d['rec'])} receptor x {len(d['lig'])} ligand atoms", "pair_evals": pairs,
                     "kernel_ms": kms, "pair_evals_per_s": pairs / (kms * 1e-3), "e2e_ms_host_buffers": wall * 1e3,
                     "poses_per_s_e2e": len(d["poses"]) / wall,
                     "roofline": {"bound": "fp32 issue", "achieved": pairs / (kms * 1e-3), "peak": ceil, "unit": "pair-evals/s",
                                  "frac": pairs / (kms * 1e-3) / ceil, "issue_slots_per_pair": slots}}
        e.close()
    except Exception as ex:  # noqa: BLE001
        out["C5"] = {"error": str(ex)}
    return out


DOCK_ISSUE_SLOTS_PER_PAIR = 24.0  # inner loop of dock_score_kernel, see profiles/dock_sass_r2.txt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)   # C4 = 1,000 timed steps after 200 warm-up
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--side", type=int, default=100, help="atoms per box edge (100 -> 1,000,000 atoms)")
    ap.add_argument("--cpu-side", type=int, default=100, help="box edge of the CPU arm (100 = the full workload)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C3 / C5 blocks (profiler runs)")
    ap.add_argument("--no-steady", action="store_true", help="skip the long steady-state window (profiler runs)")
    ap.add_argument("--no-profile", action="store_true", help="experiment: no CUDA events around the kernels")
    ap.add_argument("--profile-every", type=int, default=8)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
    args = ap.parse_args()
    if os.environ.get("MOLCHANICA_MD_LIB") and not os.environ.get("MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE"):
        # the host build of tests/cpp/host_lib/ is a checker, never a thing to be measured; an nvcc-built A/B variant has
        # to be asked for explicitly
        raise SystemExit("bench.py measures molchanica_b200/libmolchanica_md.so; unset MOLCHANICA_MD_LIB "
                         "(or set MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 for an nvcc-built A/B variant)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W_ = max(args.warmup, 3)
    K = max(args.steps, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        # at least two rebuild intervals of the fluid (28 steps each), so that the share of list builds in the window is the
        # steady-state one whatever K the caller asked for (a 20-step window holds one rebuild or none: +-30 %)
        ks, ws = min(max(K, 60), 100), min(W_, 10)
        cb = cpu_arm(args.cpu_side, ks, ws)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": cb["steps"], "warmup": ws, "ms_per_step": cb["seconds"] / cb["steps"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C4 1M-atom LJ fluid (argon, rho*=0.8442, rc=2.5 sigma, skin 1 A, dt 2 fs), "
                                       "CPU restatement of the path (oracle/md_oracle.c, OpenMP), " +
                                       ("full size" if cb["same_config"] else "bounded sample scaled by atoms"),
                           "same_config": cb["same_config"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from molchanica_b200 import _lib
    from molchanica_b200 import workloads as W
    from molchanica_b200.engine import MdEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = W.lj_fluid(m=args.side)
    n = len(w["xyzq"])
    e = MdEngine(device=local)
    if world > 1:
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            e._chk(e._L.mc_comm_unique_id(uid.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(uid).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
        e._chk(e._L.mc_comm_init(e._h, uid.ctypes.data_as(C.c_void_p), rank, world))
    lo = np.asarray(w["box_lo"], np.float32)
    e.set_box(lo, lo + w["box_ext"], True)
    e.set_cutoffs(w["rc_lj"], w["rc_q"], w["skin"], w["coul_mode"], 0.35)
    e.set_lj_table(w["ljtab"])
    e.set_atoms(w["xyzq"], w["type"], w["vel"])
    if args.lanes:
        e.set_option("pair_lanes", args.lanes)
    for kv in args.opt:
        k, v = kv.split("=")
        e.set_option(k, float(v))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def all_min_int(x):
        if world == 1:
            return int(x)
        tt = torch.tensor([x], device="cuda", dtype=torch.int64)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        return int(tt.item())

    # ---- warm-up.  W steps as asked, then -- still untimed -- on until the run is in its steady state: at least three
    # list rebuilds behind it (the first one is the initial all-gather build of a decomposed run, the second the first
    # neighbour-only migration, whose NCCL channels are set up on first use; the adaptive interval of a decomposed run
    # needs two builds to climb from its cautious start).  One step per call, so that the spacing of the rebuilds is seen.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e.step(DT_PS, W_)
    warm_total = W_
    rebuild_at = []
    nb_prev = e.stats()["n_rebuilds"]
    since = None  # steps since the last observed rebuild
    for _ in range(600):
        if len(rebuild_at) >= 3:
            break
        e.step(DT_PS, 1)
        warm_total += 1
        nb = e.stats()["n_rebuilds"]
        if since is not None:
            since += 1
        if nb != nb_prev:
            rebuild_at.append(warm_total)
            nb_prev = nb
            since = 0
    interval = e.schedule()[0] if world > 1 else (rebuild_at[-1] - rebuild_at[-2] if len(rebuild_at) >= 2 else 0)
    interval = all_min_int(interval)
    # A window shorter than the rebuild interval would otherwise be timed with or without a rebuild by accident of
    # its phase (round 1: none).  Place it so that one rebuild falls inside: conservative (1 per K instead of 1 per
    # interval), never flattering.  Every rank takes the same number of steps (decomposed runs are in lock-step).
    align = 0
    if since is not None and 0 < K < interval:
        align = max(0, interval - max(K // 2, 1) - 1 - since)
        align = all_min_int(align)
        if align:
            e.step(DT_PS, align)
            warm_total += align

    # ---- value: device-resident, CUDA events on the engine's stream around all K steps ----------
    e.set_option("profiling", 0 if args.no_profile else 1)
    e.set_option("profile_every", args.profile_every)  # CUDA-event pairs around every k-th step's kernels only
    e.reset_timers()
    s0 = e.stats()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    e.step(DT_PS, K)
    barrier()
    wall = time.perf_counter() - t0
    ms = all_max(e.last_step_ms())
    s1 = e.stats()
    value = ns_per_day(K, ms * 1e-3)

    # ---- steady state: the same measurement over a window of ~10 rebuild intervals (what a long run sees)
    steady = None
    if not args.no_steady:
        ks = int(min(max(10 * max(interval, 1), 200), 400))
        barrier()
        e.step(DT_PS, ks)
        barrier()
        ms_s = all_max(e.last_step_ms())
        s2 = e.stats()
        steady = {"value": ns_per_day(ks, ms_s * 1e-3), "unit": UNIT, "steps": ks, "ms_per_step": ms_s / ks,
                  "rebuilds": int(s2["n_rebuilds"] - s1["n_rebuilds"]),
                  "rebuild_interval_steps": ks / max(int(s2["n_rebuilds"] - s1["n_rebuilds"]), 1)}
    else:
        s2 = s1
    e.set_option("profiling", 0)

    # ---- e2e: host buffers through the C ABI, one call per step, H2D + D2H inside the timing ----
    # per step: H2D of the step's external forces (the Some(forces) argument of MdState::step,
    # reference src/mol_alignment.rs:346) from pinned memory, mc_step(dt, 1, ext), D2H of the new
    # positions (what the viewer reads back, reference src/md/mod.rs:843-852) into pinned memory.
    e2e = None
    if not args.no_e2e:
        # per step: mc_step(dt, 1, ext) with the step's external forces in pinned HOST memory (H2D inside, on its own
        # stream under the force evaluation the previous call left open -- engine.cu, option defer_tail, invisible
        # through the ABI: tests/test_gpu_parity.py::test_pipelined_external_forces...), then
        # mc_snapshot_begin hands the new positions to a pinned HOST buffer (D2H inside, double-buffered so that
        # the copy of step s overlaps the kernels of step s+1 -- the Snapshot queue of the reference,
        # src/md/mod.rs:118-152); mc_snapshot_wait(s-1) before buffer reuse, all copies drained before the clock stops.
        # A decomposed rank moves its own share: the rows of the caller's external-force array that belong to the atoms
        # it owns (gathered by the engine from the pinned array, see mc_step) in, its owned atoms + ids out.
        cap = n if world == 1 else int(3 * (n // world + n // (2 * world) + 4096))
        ext = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
        pos = [torch.empty((cap, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        ids = torch.empty((cap,), dtype=torch.int32).pin_memory()
        ext_p = C.c_void_p(ext.data_ptr())
        ids_p = C.c_void_p(ids.data_ptr()) if world > 1 else None
        n_out, epoch = C.c_int64(0), C.c_int64(0)
        seen_epoch = -1
        d2h = []

        trace = {"step": 0.0, "begin": 0.0, "wait": 0.0, "ids": 0.0} if os.environ.get("MC_E2E_TRACE") else None

        def one(k):
            nonlocal seen_epoch
            if trace is not None:
                return one_traced(k)
            e.step_raw(DT_PS, 1, ext_p)
            # positions as packed float3 (Snapshot.atom_posits); a decomposed rank's ids travel only when its layout changed
            # (*layout_epoch is in/out: the epoch whose ids this loop already holds)
            epoch.value = seen_epoch
            e._chk(e._L.mc_snapshot_begin_xyz(e._h, C.c_void_p(pos[k & 1].data_ptr()), ids_p, C.byref(n_out), C.byref(epoch)))
            b = int(n_out.value) * 12
            if world > 1 and epoch.value != seen_epoch:
                seen_epoch = epoch.value
                b += int(n_out.value) * 4
            d2h.append(b)
            if k > 0:
                e._chk(e._L.mc_snapshot_wait(e._h))

        def one_traced(k):  # MC_E2E_TRACE=1: host time of the three calls of a step, summed per rank (stderr)
            nonlocal seen_epoch
            t0 = time.perf_counter()
            e.step_raw(DT_PS, 1, ext_p)
            t1 = time.perf_counter()
            epoch.value = seen_epoch
            e._chk(e._L.mc_snapshot_begin_xyz(e._h, C.c_void_p(pos[k & 1].data_ptr()), ids_p, C.byref(n_out), C.byref(epoch)))
            t2 = time.perf_counter()
            b = int(n_out.value) * 12
            if world > 1 and epoch.value != seen_epoch:
                seen_epoch = epoch.value
                b += int(n_out.value) * 4
            t3 = time.perf_counter()
            d2h.append(b)
            if k > 0:
                e._chk(e._L.mc_snapshot_wait(e._h))
            t4 = time.perf_counter()
            trace["step"] += t1 - t0; trace["begin"] += t2 - t1; trace["ids"] += t3 - t2; trace["wait"] += t4 - t3

        def run_leg():
            # untimed: through one list rebuild of this calling pattern (a decomposed rank sizes its id / staging buffers and
            # regrows what the first rebuild outgrows; every rank takes the same number of steps)
            for k in range(max(4, int(interval) + 4)):
                one(k)
            e._chk(e._L.mc_snapshot_wait(e._h))
            ke = min(max(K, 100), 300)
            sa = e.stats()
            barrier()
            t0 = time.perf_counter()
            d2h.clear()
            worst, worst_k, t_prev = 0.0, -1, t0
            for k in range(ke):
                one(k)
                t_now = time.perf_counter()
                if t_now - t_prev > worst:
                    worst, worst_k = t_now - t_prev, k
                t_prev = t_now
            e._chk(e._L.mc_snapshot_wait(e._h))
            barrier()
            te = all_max(time.perf_counter() - t0)
            sb = e.stats()
            if trace is not None:
                print(f"[e2e trace rank {rank}] per step ms: " + ", ".join(f"{k_} {v / (ke + 4) * 1e3:.4f}" for k_, v in trace.items()),
                      file=sys.stderr, flush=True)
            return {"value": ns_per_day(ke, te), "unit": UNIT, "h2d_bytes_per_step": int(e.ext_upload_bytes()),
                    "d2h_bytes_per_step": int(sum(d2h) / max(len(d2h), 1)), "steps": ke, "ms_per_step": te / ke * 1e3,
                    "rebuilds": int(sb["n_rebuilds"] - sa["n_rebuilds"]),
                    "worst_step_ms": worst * 1e3, "worst_step_index": worst_k,
                    "api": "mc_step(ctx, dt, 1, ext_forces) + mc_snapshot_begin_xyz(ctx, out_xyz, ids-on-layout-change) / "
                           "mc_snapshot_wait(ctx), pinned host buffers; bytes are per rank (d2h averaged over the steps)"}
        e.set_option("defer_tail", 0 if any(o.startswith("defer_tail=0") for o in args.opt) else 1)
        e2e = run_leg()
        e2e["api"] += "; option defer_tail = " + ("0" if any(o.startswith("defer_tail=0") for o in args.opt) else "1 (pipelined upload)")

    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (pair force), live CUDA-event average over the timed regions
    pair_ms = s2["pair_ms_sum"] / max(s2["pair_launches_timed"], 1)
    p_full = s2["n_pairs_listed"]
    n_rows = s2["n_atoms"]
    alg_bytes = 32.0 * n_rows + 20.0 * p_full
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    traffic = limiter = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "pair_force_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_launch")
        limiter = tj.get("limiter")
        traffic_src = tj.get("source")
    n_rb = max(s2["n_rebuilds"] - s0["n_rebuilds"], 1)
    rebuild_ms = s2["build_ms_sum"] / n_rb
    build_alg = 120.0 * n_rows + 4.0 * p_full
    # `frac` follows SURVEY 8d literally: ALGORITHMIC bytes (32 N + 20 P: every listed pair counted as a 16-byte gather + a
    # 4-byte index) over the kernel time.  It can exceed 1: the gathers are served on chip (shared-memory tile / L1), DRAM
    # only carries the index stream and one pass over the positions.  `dram_frac` is the physical one: measured DRAM bytes
    # per launch (ncu) over the live kernel time against the same peak.
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "pair_force", "peak_source": peak_src,
                "dram_frac": (traffic / (pair_ms * 1e-3) / 1e9 / peak) if (traffic and pair_ms > 0 and world == 1) else None,
                "limiter": limiter, "traffic_source": traffic_src,
                "frac_note": "frac = algorithmic bytes (SURVEY 8d: 32 N + 20 P_full) / kernel time / peak; it is not a DRAM "
                             "utilisation (gathers hit on-chip memory) -- dram_frac is",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": pair_ms,
                "launches_timed": s2["pair_launches_timed"],
                "launches_in_timed_region": K, "share_of_step": pair_ms * K / ms if ms > 0 else None,
                "rebuild": {"ms": rebuild_ms, "kernel": "sort + reorder + rows_plan + rows_build (tile_build.cu)", "algorithmic_bytes": build_alg,
                            "achieved": build_alg / (rebuild_ms * 1e-3) / 1e9 if rebuild_ms > 0 else None, "unit": "GB/s",
                            "frac": build_alg / (rebuild_ms * 1e-3) / 1e9 / peak if rebuild_ms > 0 else None,
                            "interval_steps": steady["rebuild_interval_steps"] if steady else interval},
                "rebuild_ms_avg": rebuild_ms,
                "integrate_ms_avg": s2["integrate_ms_sum"] / max(s2["integrate_launches_timed"], 1),
                "halo_ms_avg": s2["halo_ms_sum"] / max(s2["halos_timed"], 1) if world > 1 else None,
                "rank0_atoms_owned": int(s2["n_atoms"]), "rank0_ghosts": int(s2["n_ghosts"]),
                "list_violations": int(s2["n_list_violations"])}

    cpu = None
    per_rank = None
    if world > 1:
        # what every rank measured (CUDA events on its own stream): in the fused halo the waits for the
        # neighbours sit inside kick_drift (ack) and the boundary rows' pair kernel (ready)
        k_int, frac = e.schedule()
        mine = torch.tensor([pair_ms, roofline["integrate_ms_avg"], roofline["rebuild_ms_avg"], float(s2["n_atoms"]),
                             float(k_int), frac], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
        per_rank = {"pair_ms": [round(float(v), 5) for v in allr[:, 0]], "integrate_ms": [round(float(v), 5) for v in allr[:, 1]],
                    "rebuild_ms": [round(float(v), 4) for v in allr[:, 2]], "atoms": [int(v) for v in allr[:, 3]],
                    "rebuild_interval": int(allr[0, 4]), "last_disp_over_half_skin": round(float(allr[0, 5]), 3)}
    roofline["per_rank"] = per_rank
    halo = (("fused peer-memory push in kick_drift + flag wait in the pair kernel" if e.halo_mode()[0]
             else "nccl send/recv (" + e.halo_mode()[1] + ")") if world > 1 else None)
    e.close()
    secondary = None
    if rank == 0 and world == 1:
        if not args.no_cpu:
            cb = cpu_arm(args.cpu_side, 40, 5)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")}
        if not args.no_secondary:
            secondary = secondary_configs(peak)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C4 {n}-atom LJ fluid (argon, rho*=0.8442, rc=2.5 sigma=8.5125 A, skin 1 A, "
                                       f"dt 2 fs, PBC {w['box_ext'][0]:.1f} A), velocity Verlet, Verlet list rebuilt on "
                                       f"displacement > skin/2",
                           "atoms": n, "l2_policy": "inputs larger than L2: list+positions = "
                                                    f"{(s2['list_bytes'] + 16 * n) / 1e6:.0f} MB per step vs 126 MB L2",
                           "parallelism": f"slab-dd{world}" if world > 1 else "single-gpu",
                           "halo": halo,
                           "pair_lanes": args.lanes or 8, "engine_options": args.opt or None,
                           "warmup_steps_run": warm_total,
                           "window": (f"{K} timed steps placed so that one list rebuild falls inside (interval {interval} steps)"
                                      if align or (0 < K < interval) else f"{K} consecutive steps, rebuilds as they come")},
                "value_steady": steady,
                "e2e": e2e, "gpu_launches": int(s1["n_kernel_launches"] - s0["n_kernel_launches"]),
                "rebuilds_in_timed_region": int(s1["n_rebuilds"] - s0["n_rebuilds"]),
                "wall_ms_per_step": wall / K * 1e3, "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary,
                "clocks": clocks, "loaded_library": _lib.LIB_PATH}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
