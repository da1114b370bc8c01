#!/bin/bash
# per-kernel times + DRAM bytes of the SURVEY 8f components on one B200
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout 120 python tools/measure_8f_kernels.py 10 > gpurun_out/measure_8f_$TAG.log 2>&1; echo "plain run rc=$?"; tail -4 gpurun_out/measure_8f_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_8f_$TAG.csv python tools/measure_8f_kernels.py 30 > gpurun_out/ncu_8f_$TAG.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary_bytes.py gpurun_out/launches_8f_$TAG.csv | tee gpurun_out/kernels_8f_$TAG.txt | head -50
