#!/bin/bash
# kernel-only timing of the default build, the gather kernel and every variant library
mkdir -p gpurun_out
TAG=${1:-r2x}
{ python tools/time_pair.py; python tools/time_pair.py pair_tile=0
for V in molchanica_b200/_variants/libmolchanica_md_*.so; do [ -f $V ] && MOLCHANICA_MD_LIB=$PWD/$V python tools/time_pair.py; done; } 2>&1 | tee gpurun_out/time_pair_$TAG.txt
