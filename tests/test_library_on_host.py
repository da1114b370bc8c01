"""The WHOLE library on the CPU.  tests/cpp/host_lib/build.sh compiles every .cu of molchanica_b200/csrc
with g++ over tests/cpp/shim_fiber/cuda_runtime.h (threads of a block = fibers, warp collectives / __syncthreads /
mbarrier + bulk copy emulated, cudaMalloc = host memory poisoned with 0xFF) into libmolchanica_md_host.so; with
MOLCHANICA_MD_LIB pointing at it the GPU parity tests run unchanged -- same Python harness, same C ABI, same engine.cu
orchestration, same kernel sources, same oracle, same tolerances -- on a machine without a GPU.

This is test infrastructure (a checker of the product's sources), not a CPU fallback: the product library still refuses to
create a handle without a B200, and nothing outside tests/ knows the host build exists.  It does not replace the -m gpu
run: timing, the compiled SASS, memory-model races, cuFFT and the multi-GPU path are only seen on hardware.

The full GPU files also pass this way (test_gpu_parity.py incl. the 1,000,000-atom case: ~5 min); here a subset that fits
the CPU suite: everything small, one 23,558-atom case, the docking scan, and the four worker scripts of the components
that were written after the round-1 GPU budget was spent (bonded terms + minimiser + between-molecules energy + clash
filter, SETTLE / SHAKE / virtual sites, SPME over a plain-DFT stand-in for cuFFT, Langevin / CSVR)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host_env():
    r = subprocess.run(["bash", os.path.join(HERE, "cpp", "host_lib", "build.sh")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    so = os.path.join(HERE, "cpp", "_build", "libmolchanica_md_host.so")
    assert os.path.exists(so)
    return dict(os.environ, MOLCHANICA_MD_LIB=so, MOLCHANICA_CUFFT_LIB=os.path.join(HERE, "cpp", "_build", "libcufft_standin.so"))


def test_gpu_parity_tests_pass_on_the_host_build(host_env):
    slow = "c4 or lane_width or variants or solv23558 or full_size or erfc_real_space or lj8000"  # (the 23.5k-atom cases: minutes on fibers)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-n", "4", "-m", "gpu", "-p", "no:cacheprovider", "-k", f"not ({slow})",
                        os.path.join(HERE, "test_gpu_parity.py"), os.path.join(HERE, "test_gpu_dock.py"),
                        os.path.join(HERE, "test_gpu_md_paths.py"), os.path.join(HERE, "test_gpu_edge_cases.py")],
                       capture_output=True, text=True, cwd=ROOT, env=host_env, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, tail
    n_passed = int(r.stdout.rsplit(" passed", 1)[0].split()[-1])
    assert n_passed >= 40, tail


@pytest.mark.parametrize("worker", ["bonded", "settle", "pme", "langevin"])
def test_component_workers_pass_on_the_host_build(worker, host_env):
    """The exact checks tests/test_gpu_{bonded,settle,pme,langevin}.py make on a GPU (the same worker scripts), here against the
    host build of the library."""
    r = subprocess.run([sys.executable, os.path.join(HERE, f"{worker}_gpu_worker.py")], capture_output=True, text=True, cwd=ROOT,
                       env=host_env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert isinstance(res, dict) and res


@pytest.mark.parametrize("case,world,halo,sched", [("lj", 2, "nccl", "fixed"), ("lj", 2, "nccl", "allgather"), ("lj", 4, "nccl", "fixed"),
                                                   ("lj", 2, "fused", "fixed"), ("lj", 2, "fused", "adaptive"), ("lj", 4, "fused", "fixed"),
                                                   ("lj", 4, "fused", "adaptive")])
def test_decomposed_run_on_the_host_build(case, world, halo, sched, host_env, oracle):
    """The multi-GPU path too: one PROCESS per rank as on the GPUs (tests/dd_worker.py, unchanged), comm.cu compiled for the
    host, NCCL replaced by a shared-memory stand-in behind the same dlopen (tests/cpp/host_lib/nccl_standin.cpp).  Slab
    decomposition, ghost selection, neighbour-only migration and all-gather rebuilds, the per-step ghost refresh with
    ncclSend / ncclRecv, rank-local snapshots: same checks as tests/test_gpu_multi.py.
    halo = fused: with MC_SHIM_SHARED_HEAP=1 the stand-in's "device" memory is a shared-memory arena per process and the
    cudaIpc calls map a neighbour's arena, so the peer-memory halo itself runs between the host processes: kick_drift<true>
    stores its boundary layers into the neighbour's arrays and raises the epoch flags, the boundary blocks of the pair kernel
    spin on them (halo_sync.cuh), the adaptive interval follows the all-gathered largest displacement."""
    import tempfile

    import numpy as np
    sys.path.insert(0, HERE)
    from dd_worker import case_workload
    from util import FORCE_RTOL, energy_close, force_rel_err, trajectory_close
    env = dict(host_env, MOLCHANICA_NCCL_LIB=os.path.join(HERE, "cpp", "_build", "libnccl_standin.so"), MC_SHIM_THREADS="2",
               MC_SHIM_SHARED_HEAP="1" if halo == "fused" else "0")
    d = tempfile.mkdtemp()
    idf, out = os.path.join(d, "nccl_id"), os.path.join(d, "out.npz")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dd_worker.py"), str(r), str(world), idf, case, out, halo, sched],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT) for r in range(world)]
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    r = np.load(out)
    w, n_steps = case_workload(case, world)
    assert int(r["fused"]) == (1 if halo == "fused" else 0), str(r["why"])
    nb = oracle.neighbors(w)
    f64, scale, en = oracle.forces(w, nb, precision=64)
    assert force_rel_err(r["f0"], f64, scale).max() < FORCE_RTOL
    assert energy_close(float(r["e_pot"]), en.sum(), f64[:, 3])
    ref = oracle.md_run(w, n_steps, precision=64)
    ok, worst, sc = trajectory_close(r["x"], ref["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, sc)
    assert int(r["violations"]) == 0 and bool(r["snap_ok"])
    assert int(r["n_owned"]) < len(w["xyzq"]) and int(r["n_ghosts"]) > 0
    assert int(r["rebuilds"]) >= 2
    if sched == "adaptive":
        assert int(r["interval"]) >= 10 and int(r["rebuilds"]) >= 2 and 0.0 < float(r["disp_frac"]) < 1.0


@pytest.mark.parametrize("case,halo", [("solvl_small", "nccl"), ("solvc_small", "fused")])  # (solvb_small = the same without a thermostat)
def test_decomposed_bonded_terms_on_the_host_build(case, halo, host_env):
    """Bonded terms (+ Langevin) on a two-rank decomposed handle of the host build against the single-handle run of the same
    build (in a subprocess, so that this process does not load the library): tests/test_gpu_multi.py's check."""
    import tempfile
    env = dict(host_env, MOLCHANICA_NCCL_LIB=os.path.join(HERE, "cpp", "_build", "libnccl_standin.so"), MC_SHIM_THREADS="2",
               MC_SHIM_SHARED_HEAP="1" if halo == "fused" else "0")
    d = tempfile.mkdtemp()
    idf, out = os.path.join(d, "nccl_id"), os.path.join(d, "out.npz")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dd_worker.py"), str(r), "2", idf, case, out, halo, "fixed"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT) for r in range(2)]
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from dd_worker import case_workload\nfrom test_gpu_multi import check_bonded_decomposed\n"
            "w, n = case_workload(%r, 2)\ncheck_bonded_decomposed(np.load(%r), w, n, %r)\nprint('OK')\n"
            % (ROOT, HERE, case, out, 1 if case.startswith("solvl") else (2 if case.startswith("solvc") else 0)))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout + p.stderr


@pytest.mark.parametrize("world,halo,per_call,defer", [(2, "fused", 1, 1), (2, "fused", 3, 1), (4, "fused", 1, 1), (2, "nccl", 1, 1), (2, "fused", 1, 0)])
def test_decomposed_external_forces_on_the_host_build(world, halo, per_call, defer, host_env, oracle):
    """mc_step(dt, k, ext) on a decomposed handle: every rank uploads its 1/N block of the caller's array, the blocks are
    all-gathered over NCCL, and the call returns after its last drift (the open force evaluation, its halo wait and -- when
    the schedule says so -- its rebuild run at the start of the next call, under that call's upload).  The trajectory must be
    the oracle's, call by call with the same arrays, in both halo modes, across rebuilds, with the deferral on and off."""
    import tempfile

    import numpy as np
    sys.path.insert(0, HERE)
    from dd_worker import case_workload, ext_forces_for_call
    from util import trajectory_close
    env = dict(host_env, MOLCHANICA_NCCL_LIB=os.path.join(HERE, "cpp", "_build", "libnccl_standin.so"), MC_SHIM_THREADS="2",
               MC_SHIM_SHARED_HEAP="1" if halo == "fused" else "0", DD_EXT=str(per_call), DD_DEFER=str(defer))
    d = tempfile.mkdtemp()
    idf, out = os.path.join(d, "nccl_id"), os.path.join(d, "out.npz")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dd_worker.py"), str(r), str(world), idf, "lj", out, halo, "fixed"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT) for r in range(world)]
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    r = np.load(out)
    w, n_steps = case_workload("lj", world)
    n = len(w["xyzq"])
    cur = dict(w)
    for k in range(n_steps // per_call):
        ref = oracle.md_run(cur, per_call, precision=64, ext_force=ext_forces_for_call(n, k))
        cur = dict(cur, xyzq=ref["xyzq"], vel=ref["vel"])
    ok, worst, sc = trajectory_close(r["x"], cur["xyzq"], w["xyzq"], w["box_ext"])
    assert ok, (worst, sc)
    verr = np.abs(r["v"][:, :3] - cur["vel"][:, :3]).max(1)
    assert np.quantile(verr, 0.99) < 2e-4 * max(float(np.abs(cur["vel"][:, :3]).max()), 1.0)
    assert int(r["violations"]) == 0 and bool(r["snap_ok"]) and int(r["rebuilds"]) >= 3
    assert int(r["ext_upload_bytes"]) <= 12 * n // world + 64  # this rank moved its block only


def test_cpp_host_mirror_runs_on_the_host_build(host_env, tmp_path):
    """include/molchanica_md.hpp (the C++ mirror of the reference's MdState interface) driven by tests/cpp/host_mirror_smoke.cpp,
    its libmolchanica_md.so resolved to the host build: known two-body answers, stepping, NPT configuration, snapshots."""
    import __graft_entry__ as g
    g.build_cpp_host()
    os.symlink(host_env["MOLCHANICA_MD_LIB"], tmp_path / "libmolchanica_md.so")
    env = dict(host_env, LD_LIBRARY_PATH=str(tmp_path) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([os.path.join(HERE, "cpp", "_build", "host_mirror_smoke"), "--npt"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr


def test_product_library_still_refuses_without_a_gpu():
    """No CPU fallback in the product: the nvcc-built library must fail loudly here."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    env = {k: v for k, v in os.environ.items() if k != "MOLCHANICA_MD_LIB"}
    code = ("from molchanica_b200.engine import MdEngine, McError\n"
            "try:\n    MdEngine()\n    print('CREATED')\nexcept McError as e:\n    print('REFUSED', e)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert "REFUSED" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr
