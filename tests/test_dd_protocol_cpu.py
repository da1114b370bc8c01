"""The N > 1 path without GPUs: two gloo processes run the decomposition protocol of comm.cu (plan
from the C ABI's mc_dd_plan, all-gather rebuild, contiguous-block halo exchange) with the oracle as
the force engine, and must reproduce the single-process trajectory."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_plan_covers_the_box_once(engine_lib):
    ext = np.array([360.3, 360.3, 360.3], np.float32)
    for world in (1, 2, 4, 8):
        owned = []
        for r in range(world):
            out = np.zeros(8, np.int32)
            assert engine_lib.mc_dd_plan(ext.ctypes.data, 9.5125, r, world, out.ctypes.data) == 0
            ncz, kz0, kz1 = int(out[2]), int(out[3]), int(out[4])
            assert kz1 - kz0 >= 2 or world == 1
            assert out[5] == (kz0 - 1) % ncz and out[6] == kz1 % ncz and out[7] == (r + 1) % world
            owned += list(range(kz0, kz1))
        assert owned == list(range(ncz))            # every layer owned exactly once, in rank order
        assert world == 1 or ncz % world == 0       # equal layer counts when the box allows it
    # too few layers for the rank count -> error, not a silent bad split
    small = np.array([30.0, 30.0, 30.0], np.float32)
    out = np.zeros(8, np.int32)
    assert engine_lib.mc_dd_plan(small.ctypes.data, 9.5, 0, 2, out.ctypes.data) != 0


@pytest.mark.parametrize("m,world", [(14, 2), (22, 2), (22, 4)])
def test_decomposition_protocol_reproduces_single_process_run(m, world, oracle, engine_lib):
    """(14, 2): two cell layers per rank, the two ranks exchange whole slabs; (22, 2): four layers per rank, disjoint
    two-layer blocks to the same peer; (22, 4): a ring of four ranks with two layers each, overlapping blocks to two
    different peers.  The worker also asserts, after every build, that owner and ghost holder keep a layer in the
    same order -- the property the zero-copy halo needs."""
    d = tempfile.mkdtemp()
    out = os.path.join(d, "out.npz")
    os.environ["DD_M"] = str(m)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29533 + m + 100 * world), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dd_protocol_worker.py"), out],
                              env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    r = np.load(out)
    sys.path.insert(0, HERE)
    from dd_protocol_worker import N_STEPS, workload
    w = workload()
    ref = oracle.md_run(w, N_STEPS, precision=64)
    dx = r["x"][:, :3].astype(np.float64) - ref["xyzq"][:, :3]
    dx -= np.rint(dx / w["box_ext"]) * w["box_ext"]
    assert np.abs(dx).max() < 5e-5, np.abs(dx).max()
    assert 0 < int(r["n_owned"]) < len(w["xyzq"]) and int(r["n_ghost"]) > 0
    assert int(r["migrations"]) == N_STEPS // 3
