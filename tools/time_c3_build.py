#!/usr/bin/env python
"""List-build and pair-kernel times of dense systems (C3: 23,558 atoms, 14 A list radius; the 55,296-atom OPC water box)
with rows_build_kernel's dense configuration on and off (option rows_dense)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molchanica_b200 import workloads as W  # noqa: E402
from molchanica_b200.engine import MdEngine  # noqa: E402

out = {}
SPLITS = [int(x) for x in os.environ.get("MC_SPLITS", "0").split(",")]
for name, w in (("C3", W.solvated_c3()), ("opc55k", W.water_box_opc(m=24, L=74.6))):
  for sp in SPLITS:
    for dense in ((1, 0) if sp == 0 else (1,)):
        e = MdEngine.from_workload(w)
        e.set_option("rows_dense", dense)
        e.set_option("build_split", sp)
        e.set_option("profiling", 1)
        for _ in range(3):
            e.build_neighbors()
        e.reset_timers()
        for _ in range(10):
            e.build_neighbors()
        sb = e.stats()
        pair_ms = e.time_pair_kernel(reps=30, flush_l2=True)
        out[f"{name}_dense{dense}_split{sp}"] = dict(atoms=len(w["xyzq"]), entries=int(sb["n_pairs_listed"]), list_MB=sb["list_bytes"] / 1e6,
                                            build_ms=sb["build_ms_sum"] / 10.0, pair_ms=pair_ms, cells=list(sb["n_cells"]))
        e.close()
print(json.dumps({k: {"build_ms": round(v["build_ms"], 4), "pair_ms": round(v["pair_ms"], 4), "list_MB": round(v["list_MB"], 1)} for k, v in out.items()}, indent=0))
