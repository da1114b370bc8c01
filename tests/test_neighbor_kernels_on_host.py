"""The cell-list kernels of neighbor.cu on the CPU (tests/cpp/neighbor_kernels_host.cpp, multi-threaded stand-in for
cuda_runtime.h): wrap + cell key, reorder + cell starts + interior flag, the 27-cell sweep with its exact accept test
and the vacuum bounding-box grid, chained as engine.cu chains them.  The Verlet list they produce must equal the
oracle's index for index -- the same bit-exact bar as on the GPU -- for periodic boxes with wrapped atoms and exclusions
and for a vacuum system."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from molchanica_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def K():
    out = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libneighbor_kernels_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I", os.path.join(HERE, "cpp", "shim_mt"), "-o", so, os.path.join(HERE, "cpp", "neighbor_kernels_host.cpp"),
                        os.path.join(HERE, "cpp", "tile_build_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(so)
    L.host_neighbor_list.restype = C.c_long
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _list_on_host(K, w, tile=None):
    """tile = None: the two-pass global sweep of neighbor.cu; tile = dict(tile_cap, list_cap, split, n_sms): the production
    single-pass build of tile_build.cu, started from those capacities."""
    n = len(w["xyzq"])
    cap = 2000 * n
    orig, flags = np.zeros(n, np.int32), np.zeros(n, np.uint8)
    xs = np.zeros((n, 4), np.float32)
    cnt, start, lst = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(cap, np.uint32)
    cells, stats = np.zeros(3, np.int32), np.zeros(4, np.int32)
    es = None if w.get("excl_start") is None else np.ascontiguousarray(w["excl_start"], np.int32)
    ei = None if w.get("excl_idx") is None else np.ascontiguousarray(w["excl_idx"], np.int32)
    if es is not None and es[-1] == 0:
        es = ei = None
    r_list = np.float32(max(w["rc_lj"], w["rc_q"])) + np.float32(w["skin"])
    tot = K.host_neighbor_list(n, _p(np.ascontiguousarray(w["xyzq"], np.float32)), _p(np.ascontiguousarray(w["box_lo"], np.float32)),
                               _p(np.ascontiguousarray(w["box_ext"], np.float32)), int(w["periodic"]), C.c_float(r_list), _p(es), _p(ei),
                               _p(orig), _p(flags), _p(xs), _p(cnt), _p(start), _p(lst), C.c_long(cap), _p(cells),
                               int(tile is not None), C.c_float(max(w["rc_lj"], w["rc_q"])), (tile or {}).get("tile_cap", 0),
                               (tile or {}).get("list_cap", 0), (tile or {}).get("split", 1), (tile or {}).get("n_sms", 1), _p(stats))
    assert tot >= 0
    if tile is not None:
        tile["stats"] = dict(launches=int(stats[0]), tile_cap=int(stats[1]), max_tile=int(stats[2]), list_cap=int(stats[3]), total=int(tot))
        tile["raw"] = (cnt.copy(), start.copy(), lst[:tot].copy(), xs.copy())
    # rows back in the caller's ids, ascending (what mc_get_neighbors exports)
    rows = [None] * n
    for k in range(n):
        rows[orig[k]] = np.sort(orig[lst[start[k]:start[k] + cnt[k]]])
    o_start = np.zeros(n + 1, np.int64)
    o_start[1:] = np.cumsum([len(r) for r in rows])
    return o_start, np.concatenate(rows).astype(np.int32), flags, orig, xs, cells


@pytest.mark.parametrize("name", ["lj", "water", "globule"])
def test_list_from_the_cuda_sources_equals_the_oracle(name, K, oracle):
    if name == "lj":
        w = W.lj_fluid(m=12)
        w["xyzq"] = w["xyzq"].copy()
        w["xyzq"][::7, 0] += np.float32(w["box_ext"][0])          # atoms outside the box: the kernel wraps them
    elif name == "water":
        w = W.water_box_c1()                                        # exclusions, a 2 x 2 x 2 grid (every cell wraps)
    else:
        w = W.globule(400, seed=17)                                 # vacuum: grid from the bounding box
    start, idx, flags, orig, xs, cells = _list_on_host(K, w)
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start), "row lengths differ"
    assert np.array_equal(idx, o_idx), "neighbour indices differ"
    if name == "lj":
        assert tuple(cells) == (4, 4, 4) and 0 < (flags & 0x80).astype(bool).sum() < len(flags)   # interior and boundary cells
        ext = np.asarray(w["box_ext"], np.float32)
        assert np.all(xs[:, :3] >= 0) and np.all(xs[:, :3] < ext)   # wrapped into the box
    if name == "water":
        assert not (flags & 0x80).any()


def _workload(name):
    if name == "lj":
        w = W.lj_fluid(m=12)
        w["xyzq"] = w["xyzq"].copy()
        w["xyzq"][::7, 0] += np.float32(w["box_ext"][0])
        return w
    if name == "water":
        return W.water_box_c1()
    return W.globule(400, seed=17)


# (workload, starting capacities).  tile_cap 32 / list_cap 1024 force the grow-and-rebuild loops of engine_build_rows;
# 5184 > 200 KB / 40 B runs the single-buffered ring; split > 1 hands slices of one cell to different work items;
# n_sms = 3 launches three persistent blocks (run one after the other here: the later ones only see the end marker).
TILE_CASES = [("lj", dict(tile_cap=32, list_cap=1024, split=1, n_sms=1)),
              ("lj", dict(tile_cap=1024, list_cap=1 << 20, split=3, n_sms=3)),
              ("water", dict(tile_cap=2048, list_cap=1 << 21, split=8, n_sms=2)),
              ("water", dict(tile_cap=5184, list_cap=1 << 21, split=2, n_sms=1)),
              ("globule", dict(tile_cap=512, list_cap=4096, split=1, n_sms=2))]


@pytest.mark.parametrize("name,tile", TILE_CASES, ids=[f"{n}-{t['tile_cap']}-{t['split']}" for n, t in TILE_CASES])
def test_production_tile_build_equals_the_oracle(name, tile, K, oracle):
    """tile_build.cu (mbarrier / bulk-copy ring emulated on OS threads): same bit-exact bar, plus the properties the force
    kernel relies on -- rows padded to 8 entries and disjoint, entries inside the force cutoff in front of the skin shell."""
    w = _workload(name)
    tile = dict(tile)
    start, idx, flags, orig, xs, cells = _list_on_host(K, w, tile)
    o_start, o_idx = oracle.neighbors(w)
    assert np.array_equal(start, o_start), "row lengths differ"
    assert np.array_equal(idx, o_idx), "neighbour indices differ"
    st = tile["stats"]
    cnt, rstart, lst, xs = tile["raw"]
    assert st["tile_cap"] >= st["max_tile"] and st["tile_cap"] % 32 == 0
    # rows: 8-aligned, disjoint, and together exactly the claimed space
    assert np.all(rstart[cnt > 0] % 8 == 0)
    padded = (cnt.astype(np.int64) + 7) & ~7
    order = np.argsort(rstart[cnt > 0], kind="stable")
    s_sorted, p_sorted = rstart[cnt > 0][order].astype(np.int64), padded[cnt > 0][order]
    assert np.all(s_sorted[1:] >= s_sorted[:-1] + p_sorted[:-1])
    assert padded.sum() == st["total"]
    # inner entries first: along a row, r2 < rc2 never follows r2 >= rc2
    rc2 = np.float32(max(w["rc_lj"], w["rc_q"])) ** 2
    ext = np.asarray(w["box_ext"], np.float64)
    n_rows_with_both = 0
    for k in np.flatnonzero(cnt > 0)[::5]:
        j = lst[rstart[k]:rstart[k] + cnt[k]]
        d = xs[j, :3].astype(np.float64) - xs[k, :3].astype(np.float64)
        if w["periodic"]:
            d -= ext * np.rint(d / ext)
        r2 = (d * d).sum(1)
        inner = r2 < float(rc2) * (1 - 1e-6)
        outer = r2 > float(rc2) * (1 + 1e-6)
        if inner.any() and outer.any():
            n_rows_with_both += 1
            assert np.flatnonzero(inner).max() < np.flatnonzero(outer).min()
    assert n_rows_with_both > 0
    if tile["tile_cap"] == 32:
        assert st["launches"] >= 3 and st["tile_cap"] > 32 and st["list_cap"] > 1024   # both capacities grew
