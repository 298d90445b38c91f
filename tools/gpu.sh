#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   — rebuild the library locally, then run the command on a B200 box
set -e
cd "$(dirname "$0")/.."
make -C lhrs_bot_b200/csrc -j8 2>&1 | grep -E "error|warning: v" || true
test -f lhrs_bot_b200/lib/liblhrs_b200.so
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
