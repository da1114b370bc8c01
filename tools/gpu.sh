#!/bin/bash
# Rebuild everything that travels to the GPU box, then run a command there.
# usage: ./scripts_gpu.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
make -C molchanica_b200/csrc -j8 2>&1 | grep -iE "error|warning: variable" || true
make -C oracle -s 2>&1 | grep -v "^built" || true
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
