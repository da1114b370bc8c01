#!/usr/bin/env python
"""Design estimate for the cluster-pair force kernel, CPU only: on a molten C4-like fluid, order
the atoms as the engine would with the Morton sub-cell key, group consecutive atoms into clusters of 2 / 4 / 8 and
measure the union neighbour list of a cluster against its atoms' own lists -- the number of gathers per listed pair
and the fraction of (atom, candidate) evaluations that hit a listed pair."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def morton_order(x, lo, ext, r_list):
    nc = np.maximum(1, np.floor(ext / (r_list * 1.001 + 1e-3)).astype(int))
    u = (x - lo) / ext * nc
    k = np.minimum(np.floor(u).astype(int), nc - 1)
    s = np.minimum(((u - k) * 4).astype(int), 3)
    part = lambda v: ((v & 2) << 2) | (v & 1)
    code = part(s[:, 0]) | (part(s[:, 1]) << 1) | (part(s[:, 2]) << 2)
    cell = (k[:, 2] * nc[1] + k[:, 1]) * nc[0] + k[:, 0]
    return np.argsort(cell * 64 + code, kind="stable")


def main():
    w = W.lj_fluid(m=20)
    r = O.md_run(w, 400, precision=32)                       # melt the simple-cubic start
    w = dict(w, xyzq=r["xyzq"])
    ext = np.asarray(w["box_ext"], np.float64)
    x = np.mod(w["xyzq"][:, :3].astype(np.float64), ext)
    r_list = w["rc_lj"] + w["skin"]
    start, idx = O.neighbors(dict(w, xyzq=np.concatenate([x, w["xyzq"][:, 3:4]], 1).astype(np.float32)))
    rows = [set(idx[start[i]:start[i + 1]].tolist()) for i in range(len(x))]
    mean_row = np.mean([len(s) for s in rows])
    print(f"{len(x)} atoms, {mean_row:.1f} list entries per atom")
    orders = {"cell-major, arbitrary inside the cell": None, "cell-major, Morton sub-cell key": morton_order(x, 0.0, ext, r_list)}
    nc = np.maximum(1, np.floor(ext / (r_list * 1.001 + 1e-3)).astype(int))
    k = np.minimum(np.floor(x / ext * nc).astype(int), nc - 1)
    orders["cell-major, arbitrary inside the cell"] = np.argsort((k[:, 2] * nc[1] + k[:, 1]) * nc[0] + k[:, 0], kind="stable")
    for name, order in orders.items():
        for c in (2, 4, 8):
            union, own = 0, 0
            for g in range(0, len(order) - c + 1, c):
                members = order[g:g + c]
                u = set()
                for i in members:
                    u |= rows[i]
                u -= set(members.tolist())
                union += len(u)
                own += sum(len(rows[i] - set(members.tolist())) for i in members)
            groups = len(order) // c
            print(f"{name:42s} cluster {c}: union list {union / groups:6.1f} per cluster = {union / own:.3f} gathers per listed pair, "
                  f"{own / (union * c):.2f} of the evaluations useful")


if __name__ == "__main__":
    main()
