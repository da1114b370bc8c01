#!/bin/bash
# Test infrastructure: builds the WHOLE library for the host (no GPU, no nvcc): every .cu of molchanica_b200/csrc is
# compiled by g++ over tests/cpp/shim_fiber/cuda_runtime.h and linked with the fiber runtime into
# tests/cpp/_build/libmolchanica_md_host.so; a plain-DFT stand-in for cuFFT and a shared-memory stand-in for NCCL go next to it.
# MOLCHANICA_MD_LIB=<that .so> makes molchanica_b200/_lib.py load it (tests/test_library_on_host.py).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../../.." && pwd)"
OUT="$ROOT/tests/cpp/_build"
OBJ="$OUT/host_lib_obj"
NAME=libmolchanica_md_host
SAN=""
if [ "$1" = "--asan" ]; then
  # AddressSanitizer build: "device" memory is heap memory, so every out-of-bounds access of a kernel or of engine.cu is
  # reported (the compute-sanitizer memcheck of this stand-in).  Load with LD_PRELOAD=$(g++ -print-file-name=libasan.so)
  # ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
  OBJ="$OUT/host_lib_obj_asan"; NAME=libmolchanica_md_host_asan; SAN="-fsanitize=address -fno-omit-frame-pointer"
fi
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
mkdir -p "$OBJ"
FLAGS="-O1 -g -std=c++20 -pthread -fPIC -ffp-contract=off -Wno-unknown-pragmas -Wno-attributes -DMC_HOST_SHIM=1 -I$ROOT/tests/cpp/shim_fiber $SAN"
SRCS="sort_scan neighbor tile_build pair_force pair_tile md_fused integrate thermostat dock dock_filter dock_poses bonded settle pme pme_params group_energy engine comm"
pids=()
for s in $SRCS; do
  src="$ROOT/molchanica_b200/csrc/$s.cu"
  if [ ! -f "$OBJ/$s.o" ] || [ -n "$(find "$ROOT/molchanica_b200/csrc" "$ROOT/tests/cpp/shim_fiber" "$ROOT/include" -newer "$OBJ/$s.o" -type f | head -1)" ]; then
    $CXX $FLAGS -x c++ -c "$src" -o "$OBJ/$s.o" &
    pids+=($!)
  fi
done
$CXX $FLAGS -c "$ROOT/tests/cpp/shim_fiber/runtime.cpp" -o "$OBJ/runtime.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
objs=""
for s in $SRCS runtime; do objs="$objs $OBJ/$s.o"; done
$CXX -shared -pthread $SAN -o "$OUT/$NAME.so" $objs -ldl -lrt
$CXX -O2 -std=c++17 -fPIC -shared -o "$OUT/libcufft_standin.so" "$HERE/cufft_standin.cpp"
$CXX -O2 -std=c++20 -fPIC -shared -pthread -o "$OUT/libnccl_standin.so" "$HERE/nccl_standin.cpp" -lrt
echo "$OUT/$NAME.so"
