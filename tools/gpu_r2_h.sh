#!/bin/bash
# ncu --set full of the list-build kernel (one launch after the melt) and of the docking scan kernel
mkdir -p gpurun_out
TAG=${1:-r2x}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_build_kernel -s 4 -c 1 \
    -o gpurun_out/tile_build_$TAG -f python bench.py --steps 40 --warmup 150 --no-cpu --no-e2e --no-secondary --no-steady > gpurun_out/ncu_tb_$TAG.log 2>&1
echo "ncu tile_build rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dock_score_kernel -s 1 -c 1 \
    -o gpurun_out/dock_$TAG -f python -c "
import sys; sys.path.insert(0,'.')
from molchanica_b200 import workloads as W
from molchanica_b200.engine import MdEngine
d=W.docking_c5(); e=MdEngine(); e.dock_score(d); e.dock_score(d); e.dock_score(d)
" > gpurun_out/ncu_dock_$TAG.log 2>&1
echo "ncu dock rc=$?"
