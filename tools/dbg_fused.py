import sys; sys.path.insert(0,'.')
import numpy as np
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
from oracle import oracle_py as O
sys.path.insert(0,'tests')
from util import trajectory_close
w=W.lj_fluid(m=12)
ref=O.md_run(w,20,precision=64)
for tile in (0,1):
  for pre in ("none","build","forces"):
    for brute in (1,0):
        e=MdEngine.from_workload(w)
        e.set_option("pair_tile",tile); e.set_option("fused_brute",brute)
        if pre!="none": e.build_neighbors()
        if pre=="forces": e.compute_forces()
        e.step(w["dt"],20)
        ok,worst,scale=trajectory_close(e.positions(),ref["xyzq"],w["xyzq"],w["box_ext"])
        dx=e.positions()[:,:3]-ref["xyzq"][:,:3]; dx-=np.rint(dx/w["box_ext"])*w["box_ext"]
        print("tile",tile,"pre",pre,"brute",brute,"ok",ok,"worst",worst,"n_bad",int((np.abs(dx).max(1)>1e-3).sum()), e.stats()["n_rebuilds"])
        e.close()
