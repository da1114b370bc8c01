"""The integrator kernels of integrate.cu on the CPU (tests/cpp/integrate_kernels_host.cpp, multi-threaded stand-in for
cuda_runtime.h): v += (F + F_ext)/m * kick * 418.4, x += v * drift; static atoms; external forces looked up by the
caller's atom id; the displacement flag with its look-ahead margin; the non-finite guard; the {tag, flag} word the last
block publishes for the host; the original-order gather / scatter.  Checked against plain numpy (the oracle's kick and
drift are the same two lines, oracle/md_oracle.c orc_kick / orc_drift)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def K():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libintegrate_kernels_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim_mt"), "-o", so, os.path.join(HERE, "cpp", "integrate_kernels_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return C.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _system(n=1500, seed=0):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.uniform(0, 40, (n, 3)), rng.normal(0, 1, (n, 1))], 1).astype(np.float32)
    v = np.concatenate([rng.normal(0, 3, (n, 3)), 1.0 / rng.uniform(1, 40, (n, 1))], 1).astype(np.float32)
    f = np.concatenate([rng.normal(0, 20, (n, 3)), rng.normal(0, 1, (n, 1))], 1).astype(np.float32)
    orig = rng.permutation(n).astype(np.int32)
    flags = np.zeros(n, np.uint8)
    flags[::9] = 1
    flags[1::9] |= 0x80                       # the engine's interior bit must not be mistaken for "static"
    ext_f = rng.normal(0, 5, (n, 3)).astype(np.float32)
    return x, v, f, orig, flags, ext_f


def test_kick_drift_matches_the_two_lines_of_the_oracle(K):
    x, v, f, orig, flags, ext_f = _system()
    n = len(x)
    kick, drift = np.float32(0.002), np.float32(0.002)
    x1, v1 = x.copy(), v.copy()
    rebuild, host = np.zeros(4, np.int32), np.zeros(2, np.int32)
    K.host_kick_drift(n, _p(x1), _p(v1), _p(f), _p(ext_f), _p(orig), _p(flags), _p(x.copy()), C.c_float(kick), C.c_float(drift),
                      C.c_float(0.5), C.c_float(1.5), _p(rebuild), _p(host), 77)
    mob = (flags & 1) == 0
    ft = f[:, :3] + ext_f[orig]                                         # external forces come in the caller's order
    s = (v[:, 3] * kick * np.float32(418.4)).astype(np.float32)
    want_v = v[:, :3] + ft * s[:, None]
    want_x = x[:, :3] + want_v * drift
    assert np.allclose(v1[mob, :3], want_v[mob], rtol=2e-6, atol=1e-6) and np.allclose(x1[mob, :3], want_x[mob], rtol=0, atol=2e-6)
    assert np.array_equal(x1[~mob], x[~mob]) and np.array_equal(v1[~mob], v[~mob])        # static atoms: untouched
    assert np.array_equal(x1[:, 3], x[:, 3]) and np.array_equal(v1[:, 3], v[:, 3])        # charge and 1/m ride along
    assert rebuild[0] == 0 and host[0] == (77 << 2)                                        # nothing moved far: tag, no flag bits


def test_displacement_flag_lookahead_and_blow_up_guard(K):
    x, v, f, orig, flags, ext_f = _system(n=700, seed=3)
    n = len(x)
    f[:] = 0
    v[:, :3] = 0
    flags[:] = 0
    xref = x.copy()
    dt = np.float32(0.002)

    def run(v_in, xref_in, max_disp, lookahead, x_in=None):
        xx, vv = (x if x_in is None else x_in).copy(), v_in.copy()
        rebuild, host = np.zeros(4, np.int32), np.zeros(2, np.int32)
        K.host_kick_drift(n, _p(xx), _p(vv), _p(f), None, _p(orig), _p(flags), _p(xref_in), C.c_float(dt), C.c_float(dt), C.c_float(max_disp),
                          C.c_float(lookahead), _p(rebuild), _p(host), 5)
        return int(rebuild[0]), int(host[0])
    # one atom displaced by 0.4 A since the build: below skin/2 = 0.5, above 0.3
    xr = xref.copy()
    xr[123, 0] -= np.float32(0.4)
    assert run(v, xr, 0.5, 0.0) == (0, 5 << 2)
    assert run(v, xr, 0.3, 0.0) == (1, (5 << 2) | 1)
    # the look-ahead: that atom also moves at 40 A/ps -> 0.08 A per drift; 1.5 drifts of margin turn 0.5 into 0.38 < 0.48
    vf = v.copy()
    vf[123, 0] = 40.0
    assert run(vf, xr, 0.5, 0.0)[0] == 0 and run(vf, xr, 0.5, 1.5)[0] == 1
    # non-finite coordinates raise bit 1
    xb = x.copy()
    xb[5, 1] = np.inf
    assert run(v, xref, 0.5, 0.0, x_in=xb)[0] & 2


def test_gather_scatter_round_trip(K):
    x, v, f, orig, flags, ext_f = _system(n=900, seed=5)
    n = len(x)
    out, back = np.zeros_like(x), np.full_like(x, 9.0)
    K.host_gather_scatter(n, _p(x), _p(orig), _p(out), _p(back), 0)
    assert np.array_equal(out[orig], x) and np.array_equal(back, x)
    back2 = np.full_like(x, 9.0)
    K.host_gather_scatter(n, _p(x), _p(orig), _p(out), _p(back2), 1)
    assert np.array_equal(back2[:, :3], x[:, :3]) and np.all(back2[:, 3] == 9.0)          # keep_w leaves the fourth component
