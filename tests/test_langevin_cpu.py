"""Langevin thermostat without a GPU.  The generator and the O step the GPU kernel runs
(molchanica_b200/csrc/langevin_terms.h, compiled for the host into a TEST library) and the oracle's independent
implementation reproduce the published Philox4x32-10 known-answer vectors, draw the same standard normals, and
the oracle's Langevin MD brings a cold fluid to the target temperature."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key) -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.fixture(scope="module")
def host_math():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liblangevin_math_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                        os.path.join(HERE, "cpp", "langevin_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_philox_known_answers_device_header_and_oracle(host_math, oracle):
    L = oracle.lib()
    for ctr, key, want in KAT:
        c, k = np.array(ctr, np.uint32), np.array(key, np.uint32)
        o1, o2 = np.zeros(4, np.uint32), np.zeros(4, np.uint32)
        host_math.lgv_host_philox(_p(c), _p(k), _p(o1))
        L.orc_philox4x32_10(_p(c), _p(k), _p(o2))
        assert tuple(int(v) for v in o1) == want and tuple(int(v) for v in o2) == want


def test_normals_are_standard_and_shared_with_the_oracle(host_math, oracle):
    n = 200000
    xi = np.zeros((n, 3), np.float32)
    host_math.lgv_host_normals(C.c_uint64(0x1234567890abcdef), C.c_int64(n), C.c_uint64(7), _p(xi))
    x = xi.astype(np.float64)
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0) < 1e-2
    assert abs((x ** 4).mean() - 3.0) < 0.06                       # kurtosis of a normal
    assert np.abs(np.corrcoef(x.T) - np.eye(3)).max() < 1e-2       # the three components are independent
    L = oracle.lib()
    for atom in (0, 1, 77777, n - 1):
        d = np.zeros(3, np.float64)
        L.orc_langevin_normals(C.c_uint64(0x1234567890abcdef), C.c_uint32(atom), C.c_uint64(7), _p(d))
        assert np.abs(d - x[atom]).max() < 2e-6
    # another step or seed gives unrelated numbers
    xj = np.zeros((n, 3), np.float32)
    host_math.lgv_host_normals(C.c_uint64(0x1234567890abcdef), C.c_int64(n), C.c_uint64(8), _p(xj))
    assert abs(np.corrcoef(xi[:, 0], xj[:, 0])[0, 1]) < 1e-2


def test_ou_step_body_matches_the_oracle_and_thermalises(host_math, oracle):
    # the kernel body on host arrays against the formula in fp64 with the oracle's normals
    rng = np.random.default_rng(2)
    n = 4096
    vel = np.concatenate([rng.normal(0, 3, (n, 3)), 1.0 / rng.uniform(1, 40, (n, 1))], 1).astype(np.float32)
    vel[::50, 3] = 0.0                                             # static atoms: untouched
    ids = rng.permutation(n).astype(np.int32)                      # the engine's internal order is not the caller's
    kT, gamma, dt = 0.0019872041 * 300.0, 5.0, 0.002
    c1 = np.exp(-gamma * dt)
    v0 = vel.copy()
    host_math.lgv_host_ou(C.c_int64(n), _p(vel), _p(ids), C.c_float(c1), C.c_float(np.sqrt(1 - c1 * c1)), C.c_float(kT),
                          C.c_uint64(99), C.c_uint64(3))
    L = oracle.lib()
    for i in (0, 1, 50, 1234):
        if v0[i, 3] == 0:
            assert np.array_equal(vel[i], v0[i])
            continue
        d = np.zeros(3, np.float64)
        L.orc_langevin_normals(C.c_uint64(99), C.c_uint32(int(ids[i])), C.c_uint64(3), _p(d))
        want = c1 * v0[i, :3].astype(np.float64) + np.sqrt(1 - c1 * c1) * np.sqrt(kT * float(v0[i, 3]) * 418.4) * d
        assert np.abs(vel[i, :3] - want).max() < 1e-5 * max(1.0, np.abs(want).max())
    # oracle MD: an LJ fluid started at 40 K is driven to 120 K
    w = W.lj_fluid(m=8, temp_k=40.0)
    r = oracle.md_run(w, 400, precision=32, want_energies=True, langevin=(120.0, 20.0, 5))
    ke = r["energies"][:, 3]
    n_at = len(w["xyzq"])
    temp = 2 * ke / (3 * n_at * 0.0019872041)
    assert temp[0] < 45 and abs(temp[250:].mean() - 120.0) < 12.0, (temp[0], temp[250:].mean())
