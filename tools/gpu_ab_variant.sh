#!/bin/bash
# A/B of a compile-time knob on one B200: builds a second copy of the library with the given nvcc flags and runs the
# parity tests of the touched path plus bench.py on both.  Example (the knob of tile_build.cu):
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab_variant.sh tile2 -DMC_TILE_MIN_BLOCKS=2'
# Output: gpurun_out/ab_<tag>_{base,variant}.json (bench lines; compare rebuild_ms_avg, ms_per_step) and the test logs.
set -u
TAG=$1; shift
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/gpurun_out" "$ROOT/molchanica_b200/_variants"
VLIB="$ROOT/molchanica_b200/_variants/libmolchanica_md_$TAG.so"
make -C "$ROOT/molchanica_b200/csrc" -j8 OUT="$VLIB" OBJDIR="$ROOT/molchanica_b200/_variants/_obj_$TAG" EXTRA_NVFLAGS="$*" > "$ROOT/gpurun_out/ab_${TAG}_build.log" 2>&1 || { tail -20 "$ROOT/gpurun_out/ab_${TAG}_build.log"; exit 1; }
cd "$ROOT"
MOLCHANICA_MD_LIB="$VLIB" timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/ab_${TAG}_tests.log 2>&1; echo "variant tests rc=$?"; tail -3 gpurun_out/ab_${TAG}_tests.log
timeout 600 python bench.py --steps 600 --warmup 100 --no-cpu > gpurun_out/ab_${TAG}_base.json 2> gpurun_out/ab_${TAG}_base.err; tail -1 gpurun_out/ab_${TAG}_base.json
MOLCHANICA_BENCH_ALLOW_LIB_OVERRIDE=1 MOLCHANICA_MD_LIB="$VLIB" timeout 600 python bench.py --steps 600 --warmup 100 --no-cpu > gpurun_out/ab_${TAG}_variant.json 2> gpurun_out/ab_${TAG}_variant.err; tail -1 gpurun_out/ab_${TAG}_variant.json
