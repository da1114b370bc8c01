"""The pose-energy scan kernel of dock.cu on the CPU (tests/cpp/dock_kernel_host.cpp, multi-threaded stand-in for
cuda_runtime.h; the compiled SASS of dock.cu is byte-identical to the GPU-validated build) against the fp64 oracle and
the committed golden fixture, at the bar of tests/test_gpu_dock.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "md_small.npz")
RTOL = 1e-5


@pytest.fixture(scope="module")
def K():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libdock_kernel_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim_mt"), "-o", so, os.path.join(HERE, "cpp", "dock_kernel_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _score(K, d, poses=None):
    """What mc_dock_score does on the host side (engine.cu), then the kernel."""
    poses = np.ascontiguousarray(d["poses"] if poses is None else poses, np.float32)
    rec, lig = np.ascontiguousarray(d["rec"], np.float32), np.ascontiguousarray(d["lig"], np.float32)
    rm = (np.asarray(d["rec_type"], np.uint32) | (np.asarray(d["rec_hphob"], np.uint32) << 16)).astype(np.uint32)
    lm = (np.asarray(d["lig_type"], np.uint32) | (np.asarray(d["lig_hphob"], np.uint32) << 16)).astype(np.uint32)
    tab = np.asarray(d["ljtab"], np.float32)
    dev_tab = np.ascontiguousarray(np.stack([tab[..., 0] ** 2, 4.0 * tab[..., 1]], -1).astype(np.float32))
    out = np.zeros((len(poses), 5), np.float32)
    K.host_dock_score(len(rec), _p(rec), _p(rm), len(lig), _p(lig), _p(lm), _p(np.ascontiguousarray(d["lig_anchor"], np.float32)),
                      tab.shape[0], tab.shape[1], _p(dev_tab), len(poses), _p(poses), _p(out))
    return out


def _check(got, ref, ref_abs):
    assert np.all(np.abs(got[:, 1] - ref[:, 1]) <= RTOL * ref_abs[:, 0] + 1e-6)
    assert np.all(np.abs(got[:, 2] - ref[:, 2]) <= 1e-5 * np.abs(ref[:, 2]) + 1e-5)
    assert np.all(np.abs(got[:, 3] - ref[:, 3]) <= RTOL * ref_abs[:, 1] + 1e-6)
    assert np.all(np.abs(got[:, 4] - ref[:, 4]) <= RTOL * ref_abs[:, 2] + 1e-6)
    score_ref = ref[:, 1].astype(np.float64) + ref[:, 2] + 10.0 * ref[:, 3]
    scale = ref_abs[:, 0] + 10.0 * ref_abs[:, 1] + np.abs(ref[:, 2])
    assert np.all(np.abs(got[:, 0] - score_ref) <= 2 * RTOL * scale + 1e-5)


def test_scan_kernel_matches_oracle_and_golden(K, oracle):
    d = W.docking_c5(n_rec=700, n_lig=20, n_poses=48, seeds=(525, 526, 527))
    ref, ref_abs = oracle.dock_score(d, precision=64, with_abs=True)
    _check(_score(K, d), ref, ref_abs)
    g = np.load(GOLD)
    gd = {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith("dock.")}
    n = min(len(gd["poses"]), 40)
    _check(_score(K, gd, poses=gd["poses"][:n]), gd["scores64"][:n], gd["abs64"][:n])
