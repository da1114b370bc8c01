"""The device prefix scan and radix sort of sort_scan.cu ("cell-list build as on-device radix-sort +
prefix-scan") on the CPU: kernels and host drivers compiled unchanged through tests/cpp/shim_mt (warp votes, shuffles,
__match_any_sync emulated on OS threads), checked bit-exactly against numpy.  The GPU parity tests cover the same code
on hardware (tests/test_gpu_parity.py::test_radix_sort_and_scan); this keeps it pinned where no GPU is present."""
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
    so = os.path.join(out, "libsort_scan_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim_mt"), "-o", so, os.path.join(HERE, "cpp", "sort_scan_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = C.CDLL(so)
    lib.host_radix_sort_pairs.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scan(K, v, align8, in_place=False):
    n = len(v)
    src = np.zeros(n + 1, np.uint32)
    src[:n] = v
    out = src if in_place else np.full(n + 1, 0xDEADBEEF, np.uint32)
    launches = C.c_int64(0)
    K.host_exclusive_scan(_p(src), _p(out), C.c_size_t(n), int(align8), C.byref(launches))
    return out, launches.value


# sizes around the tile (256 threads x 16 items = 4096) and the 1024-wide spine
@pytest.mark.parametrize("n", [0, 1, 31, 4095, 4096, 4097, 3 * 4096 + 17])
@pytest.mark.parametrize("align8", [False, True])
def test_exclusive_scan(K, n, align8):
    rng = np.random.default_rng(n + 7 * align8)
    v = rng.integers(0, 200, n).astype(np.uint32)
    out, launches = _scan(K, v, align8)
    w = ((v + 7) & ~np.uint32(7)) if align8 else v
    ref = np.concatenate([[0], np.cumsum(w.astype(np.uint64))]).astype(np.uint32)
    assert np.array_equal(out, ref)
    assert launches == (3 if n else 0)


def test_exclusive_scan_in_place_and_wraparound(K):
    """in == out (as the radix sort scans its histogram table) and 32-bit wrap-around of the running sum."""
    rng = np.random.default_rng(5)
    v = rng.integers(0, 2 ** 31, 9000, dtype=np.uint64).astype(np.uint32)
    ref = np.concatenate([np.zeros(1, np.uint64), np.cumsum(v.astype(np.uint64))]) & np.uint64(0xFFFFFFFF)
    out, _ = _scan(K, v, False, in_place=True)
    assert np.array_equal(out, ref.astype(np.uint32))


def _sort(K, keys, bits):
    n = len(keys)
    k0 = np.ascontiguousarray(keys, np.uint32).copy()
    v0 = np.arange(n, dtype=np.uint32)
    k1, v1 = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
    launches = C.c_int64(0)
    cur = K.host_radix_sort_pairs(_p(k0), _p(v0), _p(k1), _p(v1), C.c_size_t(n), bits, C.byref(launches))
    return (k1[:n], v1[:n]) if cur else (k0, v0), launches.value


# one warp tile is 512 keys, one block 4096: ragged tails exercise the partial-warp vote masks
@pytest.mark.parametrize("n,bits", [(0, 16), (1, 8), (33, 8), (511, 16), (512, 16), (513, 16), (4096 + 77, 24), (10000, 32)])
def test_radix_sort_is_the_stable_sort(K, n, bits):
    rng = np.random.default_rng(n * 31 + bits)
    keys = rng.integers(0, 2 ** bits, n, dtype=np.uint64).astype(np.uint32)
    (ks, vs), launches = _sort(K, keys, bits)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(vs, order) and np.array_equal(ks, keys[order])
    assert launches == (5 * ((bits + 7) // 8) if n else 0)


def test_radix_sort_few_distinct_keys_keeps_input_order(K):
    """Cell keys: many atoms per key.  Stability is what makes the within-cell order identical on every rank that sorts the
    same candidates (the zero-copy halo of comm.cu relies on it)."""
    rng = np.random.default_rng(11)
    keys = rng.integers(0, 7, 6000).astype(np.uint32) * 37
    (ks, vs), _ = _sort(K, keys, 16)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(vs, order)
    same = ks[1:] == ks[:-1]
    assert np.all(vs[1:][same] > vs[:-1][same])


def test_radix_sort_ignores_bits_above_the_requested_width(K):
    """radix_sort_pairs(bits) orders by the low `bits` only: the cell-list build passes the width of the largest key."""
    rng = np.random.default_rng(12)
    keys = rng.integers(0, 2 ** 20, 3000, dtype=np.uint64).astype(np.uint32)
    (ks, vs), _ = _sort(K, keys, 8)
    order = np.argsort(keys & 0xFF, kind="stable").astype(np.uint32)
    assert np.array_equal(vs, order)
