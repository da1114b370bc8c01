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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        note = None
        if not self.rows:
            # the sampler produced nothing (interval refused, process too slow to start): one query right after the region
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=20).stdout
                self.rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
                note = "sampler returned nothing; single query right after the timed region"
            except Exception:
                pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_arm(side, steps, warmup):
    """The CPU path (oracle port, OpenMP over all host cores) on a bounded sample: the same
    fluid at the same density in a side^3-atom box; ns/day is reported for the 1M-atom system by
    scaling with atoms (the work per atom is identical)."""
    from molchanica_b200 import workloads as W
    from oracle import oracle_py as O
    O.lib()
    w = W.lj_fluid(m=side)
    n = len(w["xyzq"])
    st = O.md_run(w, warmup, precision=32)
    t0 = time.perf_counter()
    O.md_run(w, steps, precision=32, xyzq=st["xyzq"], vel=st["vel"])
    dt = time.perf_counter() - t0
    v_sample = ns_per_day(steps, dt)
    v_1m = v_sample * n / 1.0e6
    return dict(value=v_1m, unit=UNIT, cores=O.num_threads(), kind="port",
                sample=f"{n}-atom box of the same LJ fluid (side {side}), {steps} steps incl. list rebuilds, "
                       f"{dt:.2f} s; scaled by atoms to 1M",
                seconds=dt, steps=steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)   # C4 = 1,000 timed steps after 200 warm-up
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--side", type=int, default=100, help="atoms per box edge (100 -> 1,000,000 atoms)")
    ap.add_argument("--cpu-side", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler runs)")
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
        cb = cpu_arm(args.cpu_side, min(K, 200), min(W_, 10))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": cb["steps"], "warmup": min(W_, 10), "ms_per_step": cb["seconds"] / cb["steps"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C4 1M-atom LJ fluid (argon, rho*=0.8442, rc=2.5 sigma, skin 1 A, dt 2 fs), "
                                       "CPU restatement on a bounded sample"},
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

    # ---- value: device-resident, CUDA events on the engine's stream around all K steps ----------
    e.step(DT_PS, W_)
    e.set_option("profiling", 0 if args.no_profile else 1)
    e.set_option("profile_every", args.profile_every)  # CUDA-event pairs around every k-th step's kernels only
    e.reset_timers()
    s0 = e.stats()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    e.step(DT_PS, K)
    barrier()
    wall = time.perf_counter() - t0
    ms = e.last_step_ms()
    clocks = sampler.stop() if rank == 0 else None
    s1 = e.stats()
    e.set_option("profiling", 0)
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = ns_per_day(K, ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, one call per step, H2D + D2H inside the timing ----
    # per step: H2D of the step's external forces (the Some(forces) argument of MdState::step,
    # reference src/mol_alignment.rs:346) from pinned memory, mc_step(dt, 1, ext), D2H of the new
    # positions (what the viewer reads back, reference src/md/mod.rs:843-852) into pinned memory.
    e2e = None
    if not args.no_e2e:
        # per step: mc_step(dt, 1, ext) with the step's external forces in pinned HOST memory (H2D inside; on a single GPU
        # the upload runs on its own stream under the force evaluation the previous call left open -- engine.cu,
        # option defer_tail, invisible through the ABI: tests/newpaths_md.py::test_pipelined_external_forces...), then
        # mc_snapshot_begin hands the new positions to a pinned HOST buffer (D2H inside, double-buffered so that
        # the copy of step s overlaps the kernels of step s+1 -- the Snapshot queue of the reference,
        # src/md/mod.rs:118-152); mc_snapshot_wait(s-1) before buffer reuse, all copies drained before the clock stops.
        # A decomposed rank moves its own share: the full ext-force array in, its owned atoms + ids out.
        cap = n if world == 1 else int(3 * (n // world + n // (2 * world) + 4096))
        ext = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
        pos = [torch.empty((cap, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        ids = [torch.empty((cap,), dtype=torch.int32).pin_memory() for _ in range(2)]
        ext_p = C.c_void_p(ext.data_ptr())
        n_out = C.c_int64(0)
        d2h = 0

        def one(k):
            nonlocal d2h
            e.step_raw(DT_PS, 1, ext_p)
            e._chk(e._L.mc_snapshot_begin(e._h, C.c_void_p(pos[k & 1].data_ptr()),
                                          C.c_void_p(ids[k & 1].data_ptr()) if world > 1 else None, C.byref(n_out)))
            d2h = int(n_out.value) * (16 + (4 if world > 1 else 0))
            if k > 0:
                e._chk(e._L.mc_snapshot_wait(e._h))
        def run_leg():
            for k in range(4):
                one(k)
            e._chk(e._L.mc_snapshot_wait(e._h))
            ke = min(K, 300)
            barrier()
            t0 = time.perf_counter()
            for k in range(ke):
                one(k)
            e._chk(e._L.mc_snapshot_wait(e._h))
            barrier()
            te = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([te], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                te = float(tt.item())
            return {"value": ns_per_day(ke, te), "unit": UNIT, "h2d_bytes_per_step": int(ext.numel() * 4),
                    "d2h_bytes_per_step": d2h, "steps": ke, "ms_per_step": te / ke * 1e3,
                    "api": "mc_step(ctx, dt, 1, ext_forces) + mc_snapshot_begin(ctx, out, ids) / mc_snapshot_wait(ctx), pinned host "
                           "buffers; bytes are per rank"}
        try:
            if world == 1:
                e.set_option("defer_tail", 0 if any(o.startswith("defer_tail=0") for o in args.opt) else 1)
            e2e = run_leg()
            if world == 1:
                e2e["api"] += "; option defer_tail = " + ("0" if any(o.startswith("defer_tail=0") for o in args.opt) else "1 (pipelined upload)")
        except Exception as ex:  # noqa: BLE001
            if world > 1:
                raise
            # the pipelined upload of a single-GPU handle (option defer_tail) was written after this round's last hardware
            # run: should it fail here, the line still carries the path that WAS measured, and says so
            try:
                e.set_option("defer_tail", 0)
                e2e = run_leg()
                e2e["note"] = f"pipelined upload failed ({ex}); measured with defer_tail = 0"
            except Exception as ex2:  # noqa: BLE001
                e2e = {"value": None, "unit": UNIT, "error": f"{ex}; then {ex2}"}

    # ---- roofline of the dominant kernel (pair force), live CUDA-event average over the timed region
    pair_ms = (s1["pair_ms_sum"] - 0.0) / max(s1["pair_launches_timed"], 1)
    p_full = s1["n_pairs_listed"]
    n_rows = s1["n_atoms"]
    alg_bytes = 32.0 * n_rows + 20.0 * p_full
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "pair_force_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "pair_force_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": pair_ms,
                "launches_timed": s1["pair_launches_timed"],
                "launches_in_timed_region": K, "share_of_step": pair_ms * K / ms if ms > 0 else None,
                "rebuild_ms_avg": s1["build_ms_sum"] / max(s1["n_rebuilds"] - s0["n_rebuilds"], 1),
                "integrate_ms_avg": s1["integrate_ms_sum"] / max(s1["integrate_launches_timed"], 1),
                "halo_ms_avg": s1["halo_ms_sum"] / max(s1["halos_timed"], 1) if world > 1 else None,
                "rank0_atoms_owned": int(s1["n_atoms"]), "rank0_ghosts": int(s1["n_ghosts"]),

                "list_violations": int(s1["n_list_violations"])}

    cpu = None
    per_rank = None
    if world > 1:
        # what every rank measured (CUDA events on its own stream): in the fused halo the waits for the
        # neighbours sit inside kick_drift (ack) and the boundary rows' pair kernel (ready)
        k_int, frac = e.schedule()
        mine = torch.tensor([pair_ms, roofline["integrate_ms_avg"], roofline["rebuild_ms_avg"], float(s1["n_atoms"]),
                             float(k_int), frac], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
        per_rank = {"pair_ms": [round(float(v), 5) for v in allr[:, 0]], "integrate_ms": [round(float(v), 5) for v in allr[:, 1]],
                    "rebuild_ms": [round(float(v), 4) for v in allr[:, 2]], "atoms": [int(v) for v in allr[:, 3]],
                    "rebuild_interval": int(allr[0, 4]), "last_disp_over_half_skin": round(float(allr[0, 5]), 3)}
    roofline["per_rank"] = per_rank
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_arm(args.cpu_side, 100, 5)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C4 {n}-atom LJ fluid (argon, rho*=0.8442, rc=2.5 sigma=8.5125 A, skin 1 A, "
                                       f"dt 2 fs, PBC {w['box_ext'][0]:.1f} A), velocity Verlet, Verlet list rebuilt on "
                                       f"displacement > skin/2",
                           "atoms": n, "l2_policy": "inputs larger than L2: list+positions = "
                                                    f"{(4 * p_full + 16 * n) / 1e6:.0f} MB per step vs 126 MB L2",
                           "parallelism": f"slab-dd{world}" if world > 1 else "single-gpu",
                           "halo": (("fused peer-memory push in kick_drift + flag wait in the pair kernel" if e.halo_mode()[0]
                                     else "nccl send/recv (" + e.halo_mode()[1] + ")") if world > 1 else None),
                           "pair_lanes": args.lanes or 8, "engine_options": args.opt or None},
                "e2e": e2e, "gpu_launches": int(s1["n_kernel_launches"] - s0["n_kernel_launches"]),
                "rebuilds_in_timed_region": int(s1["n_rebuilds"] - s0["n_rebuilds"]),
                "wall_ms_per_step": wall / K * 1e3, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "loaded_library": _lib.LIB_PATH}
        print(json.dumps(line))
    e.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
