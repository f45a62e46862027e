#!/bin/bash
# tools/gpu.sh [gpurun options] -- 'command'   — rebuild both libraries HERE first (the box only gets prebuilt .so files),
# then hand over to gpurun.  A stale libbacon_ivp.so once cost a 10-minute GPU call.
set -e
cd "$(dirname "$0")/.."
make -C bacon_b200/csrc -j8 2>&1 | grep -E "error|Error" && exit 1
make -C oracle 2>&1 | grep -E "error|Error" && exit 1
mkdir -p gpurun_out
exec /usr/local/graft/bin/gpurun "$@"
