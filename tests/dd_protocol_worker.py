"""CPU mirror of the decomposition PROTOCOL of molchanica_b200/csrc/comm.cu, one process per rank over
torch.distributed (gloo): whole-cell-layer slabs from mc_dd_plan, the fixed-capacity all-gather
first build with -1 padding, the neighbour-only migration of the later builds (last two / first two
owned layers exchanged with the ring neighbours, candidate order [from prev | own | from next]), the
stable (layer, cy, cx) sort that makes owned / boundary / ghost blocks contiguous AND identically
ordered on owner and ghost holder, and the block-for-block halo exchange with the two-rank ordering.  The force
engine of every rank is the oracle (this is test infrastructure); the test compares the decomposed
trajectory with a single-process oracle run.
usage: dd_protocol_worker.py <out_npz>   (RANK / WORLD_SIZE / MASTER_* from the environment)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import _lib  # noqa: E402
from molchanica_b200 import workloads as W  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

N_STEPS, REBUILD_EVERY = 9, 3
f32 = np.float32


def workload():
    # DD_M = 14: four cell layers, two per rank (the ranks exchange whole slabs); 22: eight layers (two-layer
    # blocks).  Hot, so that atoms do cross slab boundaries within the nine steps.
    return W.lj_fluid(m=int(os.environ.get("DD_M", "14")), temp_k=3000.0)


def main():
    out = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = workload()
    n = len(w["xyzq"])
    ext = np.asarray(w["box_ext"], f32)
    r_list = f32(max(w["rc_lj"], w["rc_q"])) + f32(w["skin"])
    plan = np.zeros(8, np.int32)
    rc = _lib.lib().mc_dd_plan(ext.ctypes.data, float(r_list), rank, world, plan.ctypes.data)
    assert rc == 0, rc
    ncx, ncy, ncz, kz0, kz1 = (int(v) for v in plan[:5])
    nl = kz1 - kz0
    prev, nxt = (rank + world - 1) % world, int(plan[7])
    assert nxt == (rank + 1) % world
    inv_cw = (np.array([ncx, ncy, ncz], np.float64) / ext.astype(np.float64)).astype(f32)
    cap = n // world + n // (2 * world) + 4096

    # initial ownership: an index block (comm_set_atoms); the first rebuild redistributes by position
    lo, hi = n * rank // world, n * (rank + 1) // world
    own = dict(x=w["xyzq"][lo:hi].copy(), v=w["vel"][lo:hi].copy(), id=np.arange(lo, hi, dtype=np.int32))
    S = {}

    def sort_into_layers(x, v, ids):
        x[:, :3] -= np.floor(x[:, :3] * (f32(1) / ext)) * ext      # wrap into the box (dd_key_kernel)
        cell = np.minimum(np.floor(x[:, :3] * inv_cw).astype(np.int64), [ncx - 1, ncy - 1, ncz - 1])
        layer = (cell[:, 2] - (kz0 - 1)) % ncz                       # local layer: 0 = ghost from prev
        keep = layer < nl + 2
        key = (layer * ncy + cell[:, 1]) * ncx + cell[:, 0]
        order = np.argsort(key[keep], kind="stable")
        S["x"], S["v"], S["id"] = x[keep][order], v[keep][order], ids[keep][order]
        lay = layer[keep][order]
        off = np.searchsorted(lay, np.arange(nl + 3))                # slot where each local layer starts
        S["o_own"], S["o_first_end"], S["o_last_begin"], S["o_own_end"], S["o_end"] = \
            int(off[1]), int(off[2]), int(off[nl]), int(off[nl + 1]), int(off[nl + 2])
        # the record every rank all-gathers (dd_layer_offsets_kernel): [1] first owned, [4] end owned,
        # [8] end of the second owned layer, [9] begin of the second-to-last owned layer
        t = np.zeros(16, np.int64)
        t[1], t[4] = off[1], off[nl + 1]
        t[8] = off[3] if nl >= 2 else off[2]
        t[9] = off[nl - 1] if nl >= 2 else off[nl]
        tabs = [torch.zeros(16, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(tabs, torch.from_numpy(t))
        S["table"] = [tt.numpy() for tt in tabs]
        assert sum(int(tt[4] - tt[1]) for tt in S["table"]) == n      # every atom has exactly one owner
        check_ghost_order()

    def check_ghost_order():
        """The property the zero-copy halo relies on: my first owned layer holds the same atoms IN THE SAME ORDER
        as prev's ghost block behind its owned slots, my last layer as next's ghost block in front."""
        ids = S["id"]
        first = torch.from_numpy(ids[S["o_own"]:S["o_first_end"]].copy())
        last = torch.from_numpy(ids[S["o_last_begin"]:S["o_own_end"]].copy())
        g_next = torch.empty((S["o_end"] - S["o_own_end"],), dtype=torch.int32)
        g_prev = torch.empty((S["o_own"],), dtype=torch.int32)
        ops = [dist.P2POp(dist.isend, first, prev, tag=0), dist.P2POp(dist.isend, last, nxt, tag=1),
               dist.P2POp(dist.irecv, g_next, nxt, tag=0), dist.P2POp(dist.irecv, g_prev, prev, tag=1)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        assert np.array_equal(g_next.numpy(), ids[S["o_own_end"]:S["o_end"]]), "ghost-next block order differs from the owner's"
        assert np.array_equal(g_prev.numpy(), ids[:S["o_own"]]), "ghost-prev block order differs from the owner's"

    def blocks(t):
        own_n = int(t[4] - t[1])
        whole = world == 2 and t[8] > t[9]
        return (own_n, own_n, 0) if whole else (int(t[8] - t[1]), int(t[9] - t[1]), int(t[4] - t[9]))

    def migrate():
        """comm_rebuild, neighbour-only: sizes come from the table of the previous build."""
        me, tp, tn = S["table"][rank], S["table"][prev], S["table"][nxt]
        my_lo, my_hi_off, my_hi = blocks(me)
        from_next, from_prev = blocks(tn)[0], blocks(tp)[2]
        o = owned()
        n_own = len(o["id"])
        whole = world == 2 and my_hi == 0 and from_prev == 0
        off_own = (0 if rank == 0 else from_next) if whole else from_prev
        off_next = (n_own if rank == 0 else 0) if whole else from_prev + n_own
        n_all = n_own + from_next + from_prev
        arr = {"x": (np.zeros((n_all, 4), f32), torch.float32), "v": (np.zeros((n_all, 4), f32), torch.float32),
               "id": (np.full((n_all,), -1, np.int32), torch.int32)}
        ops, recvs = [], []
        for k, name in enumerate(("x", "v", "id")):
            a = arr[name][0]
            a[off_own:off_own + n_own] = o[name]
            if my_lo:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(o[name][:my_lo].copy()), prev, tag=10 + k))
            if my_hi:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(o[name][my_hi_off:my_hi_off + my_hi].copy()), nxt, tag=20 + k))
            if from_next:
                t_ = torch.empty((from_next,) + a.shape[1:], dtype=arr[name][1])
                ops.append(dist.P2POp(dist.irecv, t_, nxt, tag=10 + k))
                recvs.append((a, off_next, t_))
            if from_prev:
                t_ = torch.empty((from_prev,) + a.shape[1:], dtype=arr[name][1])
                ops.append(dist.P2POp(dist.irecv, t_, prev, tag=20 + k))
                recvs.append((a, 0, t_))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        for a, off, t_ in recvs:
            a[off:off + len(t_)] = t_.numpy()
        assert (arr["id"][0] >= 0).all() and len(np.unique(arr["id"][0])) == n_all
        sort_into_layers(arr["x"][0], arr["v"][0], arr["id"][0])

    def rebuild():
        def padded(a, fill):
            buf = np.full((cap,) + a.shape[1:], fill, a.dtype)
            buf[:len(a)] = a
            return torch.from_numpy(buf)
        gx = [torch.empty((cap, 4), dtype=torch.float32) for _ in range(world)]
        gv = [torch.empty((cap, 4), dtype=torch.float32) for _ in range(world)]
        gi = [torch.empty((cap,), dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gx, padded(own["x"], 0))
        dist.all_gather(gv, padded(own["v"], 0))
        dist.all_gather(gi, padded(own["id"], -1))
        x, v, ids = torch.cat(gx).numpy(), torch.cat(gv).numpy(), torch.cat(gi).numpy()
        live = ids >= 0
        assert live.sum() == n and len(np.unique(ids[live])) == n  # every atom owned exactly once
        x, v, ids = x[live], v[live], ids[live]
        sort_into_layers(x, v, ids)

    def halo():
        x = S["x"]
        first = torch.from_numpy(x[S["o_own"]:S["o_first_end"]].copy())
        last = torch.from_numpy(x[S["o_last_begin"]:S["o_own_end"]].copy())
        g_next = torch.empty((S["o_end"] - S["o_own_end"], 4), dtype=torch.float32)
        g_prev = torch.empty((S["o_own"], 4), dtype=torch.float32)
        # same order as comm_halo_positions: send prev, send next, recv next, recv prev
        ops = [dist.P2POp(dist.isend, first, prev, tag=0), dist.P2POp(dist.isend, last, nxt, tag=1),
               dist.P2POp(dist.irecv, g_next, nxt, tag=0), dist.P2POp(dist.irecv, g_prev, prev, tag=1)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        x[S["o_own_end"]:S["o_end"]] = g_next.numpy()
        x[:S["o_own"]] = g_prev.numpy()

    def forces():
        lw = dict(w, xyzq=S["x"], type=np.zeros(len(S["x"]), np.uint16), excl_start=None, excl_idx=None, pairs14=None)
        f, _, _ = O.forces(lw, O.neighbors(lw), precision=64)
        return f[S["o_own"]:S["o_own_end"]]

    def kick(f, half_dt):
        sl = slice(S["o_own"], S["o_own_end"])
        s = (S["v"][sl, 3] * f32(half_dt) * f32(418.4)).astype(f32)
        S["v"][sl, :3] += f[:, :3] * s[:, None]

    def owned():
        sl = slice(S["o_own"], S["o_own_end"])
        return dict(x=S["x"][sl].copy(), v=S["v"][sl].copy(), id=S["id"][sl].copy())

    dt = f32(w["dt"])
    rebuild()
    f = forces()
    for step in range(N_STEPS):
        kick(f, 0.5 * dt)
        sl = slice(S["o_own"], S["o_own_end"])
        S["x"][sl, :3] += S["v"][sl, :3] * dt
        if (step + 1) % REBUILD_EVERY == 0:
            migrate()
            S["migrated"] = S.get("migrated", 0) + 1
        else:
            halo()
        f = forces()
        kick(f, 0.5 * dt)
    # gather the final state by original id
    res = np.zeros((n, 4), np.float32)
    o = owned()
    res[o["id"]] = o["x"]
    t = torch.from_numpy(res)
    dist.all_reduce(t)
    if rank == 0:
        np.savez(out, x=t.numpy(), migrations=S.get("migrated", 0), n_owned=len(o["id"]), n_ghost=S["o_end"] - (S["o_own_end"] - S["o_own"]),
                 plan=plan)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
