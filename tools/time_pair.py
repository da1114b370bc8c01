#!/usr/bin/env python
"""Times the pair-force kernel alone on fixed positions (mc_time_kernels: L2 flushed between launches) for the C4 fluid
after a short melt; prints one JSON line.  MOLCHANICA_MD_LIB selects a variant build.  usage: time_pair.py [side] [opt=val ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
w = W.lj_fluid(m=side)
e = MdEngine.from_workload(w)
e.set_option("pair_tile", 0)   # melt with the gather kernel (an experimental variant of the tile kernel may compute nothing)
e.step(w["dt"], 150)       # molten, a few rebuilds behind
e.set_option("pair_tile", 2)
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=")
        e.set_option(k, float(v))
e.build_neighbors()
ms = e.time_pair_kernel(reps=30, flush_l2=True)
st = e.stats()
print(json.dumps({"lib": os.environ.get("MOLCHANICA_MD_LIB", "default"), "opts": [a for a in sys.argv[1:] if "=" in a], "pair_ms": ms,
                  "pairs": st["n_pairs_listed"], "list_bytes": st["list_bytes"]}))
