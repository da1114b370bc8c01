#!/bin/bash
# Compares the SASS of every kernel-bearing object of the measured path with the same files of another commit
# (default: the last commit that ran on B200s).  "SAME" = the device code is byte-identical, i.e. refactors for
# host-side testing (fenced launchers, shared headers) did not touch what was validated on hardware.
# usage: tools/sass_diff.sh [commit]   (needs nvcc; no GPU)
set -e
BASE=${1:-ffc3728}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
mkdir -p $TMP/csrc $TMP/include
FILES="sort_scan neighbor tile_build pair_force pair_tile md_fused integrate thermostat dock dock_filter bonded settle pme group_energy comm"
for f in $(git -C $ROOT ls-tree --name-only $BASE molchanica_b200/csrc/ | grep -E "\.(cu|cuh|h)$"); do
  git -C $ROOT show $BASE:$f > $TMP/csrc/$(basename $f)
done
git -C $ROOT show $BASE:include/molchanica_md.h > $TMP/include/molchanica_md.h
sed -i "s#../../include/molchanica_md.h#$TMP/include/molchanica_md.h#" $TMP/csrc/common.cuh
make -C $ROOT/molchanica_b200/csrc -j8 > /dev/null
sass() { cuobjdump -sass $1 | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{$1=$1; print}' | md5sum | cut -c1-12; }
rc=0
for f in $FILES; do
  (cd $TMP/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -c $f.cu -o $f.o)
  a=$(sass $TMP/csrc/$f.o); b=$(sass $ROOT/molchanica_b200/csrc/_obj/$f.o)
  if [ "$a" = "$b" ]; then echo "$f $a SAME"; else echo "$f $BASE=$a now=$b DIFFERENT"; rc=1; fi
done
rm -rf $TMP
exit $rc
