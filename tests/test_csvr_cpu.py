"""CSVR thermostat without a GPU: the scale factor the single device thread computes
(molchanica_b200/csrc/csvr_terms.h, compiled for the host) equals the oracle's independent implementation, has the
limits of the Bussi-Donadio-Parrinello formula, and -- applied over and over -- samples the canonical distribution of
the kinetic energy; the oracle's CSVR MD thermalises a cold fluid."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
KB = 0.0019872041


@pytest.fixture(scope="module")
def host_math():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liblangevin_math_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so,
                        os.path.join(HERE, "cpp", "langevin_math_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.csvr_host_lambda.restype = C.c_double
    L.csvr_host_lambda.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64]
    return L


def test_lambda_matches_the_oracle_and_has_the_right_limits(host_math, oracle):
    L = oracle.lib()
    kT = KB * 300.0
    for nf, k, c, seed, step in ((3.0 * 648 - 3 * 216, 380.0, 0.98, 7, 0), (2997.0, 900.0, 0.5, 1 << 40, 123456789012),
                                 (2.0, 0.5, 0.9, 3, 4), (1.0, 0.2, 0.9, 3, 5), (30000.0, 1e4, 0.999, 11, 2)):
        a = host_math.csvr_host_lambda(k, kT, nf, c, seed, step)
        b = L.orc_csvr_lambda(k, kT, nf, c, seed, step)
        assert a > 0 and abs(a - b) < 1e-12 * a, (nf, a, b)
    # c = 1 (tau -> infinity): no coupling, lambda = 1 exactly; no kinetic energy: nothing to scale
    assert host_math.csvr_host_lambda(123.0, kT, 600.0, 1.0, 5, 9) == 1.0
    assert host_math.csvr_host_lambda(0.0, kT, 600.0, 0.5, 5, 9) == 1.0
    # c = 0 (tau -> 0): the new kinetic energy is a fresh canonical sample, independent of the old one
    l1 = host_math.csvr_host_lambda(10.0, kT, 600.0, 0.0, 5, 9)
    l2 = host_math.csvr_host_lambda(1000.0, kT, 600.0, 0.0, 5, 9)
    assert abs(l1 * l1 * 10.0 - l2 * l2 * 1000.0) < 1e-9 * l1 * l1 * 10.0


@pytest.mark.parametrize("nf", [6.0, 300.0])
def test_repeated_rescaling_samples_the_canonical_kinetic_energy(nf, host_math):
    """K ~ Gamma(shape Nf/2, scale kT): mean Nf kT / 2, variance Nf kT^2 / 2."""
    kT, c = KB * 250.0, 0.6
    k = 5.0 * nf * kT                     # start far from equilibrium
    ks = []
    for step in range(24000):
        lam = host_math.csvr_host_lambda(k, kT, nf, c, 2024, step)
        k *= lam * lam
        if step >= 200:
            ks.append(k)
    ks = np.array(ks)
    mean, var = 0.5 * nf * kT, 0.5 * nf * kT * kT
    n_eff = len(ks) * (1 - c) / (1 + c)   # the chain is correlated with coefficient ~c
    assert abs(ks.mean() - mean) < 5 * np.sqrt(var / n_eff)
    assert abs(ks.var() - var) < 0.12 * var


def test_oracle_csvr_md_thermalises(oracle):
    # (the simple-cubic start of the fluid keeps releasing potential energy for the first picosecond: couple tightly)
    w = W.lj_fluid(m=8, temp_k=40.0)
    r = oracle.md_run(w, 500, precision=32, want_energies=True, csvr=(120.0, 100.0, 5))
    temp = 2 * r["energies"][:, 3] / (3 * len(w["xyzq"]) * KB)
    assert temp[0] < 45 and abs(temp[300:].mean() - 120.0) < 10.0, (temp[0], temp[300:].mean())
