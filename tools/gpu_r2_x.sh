#!/bin/bash
# dense-system list build (rows_dense on / off): parity of both list-build kernels, then C3 and the OPC box timed
mkdir -p gpurun_out
TAG=${1:-r2x}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_list_build or bit_exact" > gpurun_out/pytest_build_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_build_$TAG.log
timeout 200 python tools/time_c3_build.py > gpurun_out/c3_build_$TAG.json 2> gpurun_out/c3_build_$TAG.err; cat gpurun_out/c3_build_$TAG.json; tail -3 gpurun_out/c3_build_$TAG.err
