#!/bin/bash
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS"  ->  variants/libbacon_ivp_NAME.so
# (plain kernels only: no terminal-event instantiations)
# A/B timing of kernel variants on one GPU box:  BACON_IVP_LIB=variants/libbacon_ivp_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
B=/tmp/bacon_variant_$NAME; mkdir -p $B variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $EXTRA"
S=bacon_b200/csrc
$NV -c $S/engine.cu -o $B/engine.o &
$NV -DBACON_SKIP_BDF -DBACON_SKIP_ADAMS -DBACON_SKIP_EVENTS -Xptxas -v -c $S/rhs_builtin.cu -o $B/fast.o 2> $B/ptxas_fast.log &
$NV -DBACON_SKIP_BDF -DBACON_SKIP_ADAMS -DBACON_SKIP_EVENTS -fmad=false -DBACON_STRICT_FP -c $S/rhs_builtin.cu -o $B/strict.o &
$NV -DBACON_SKIP_RK -DBACON_SKIP_EVENTS -c $S/rhs_builtin.cu -o $B/bdf_fast.o &
$NV -DBACON_SKIP_RK -DBACON_SKIP_EVENTS -fmad=false -DBACON_STRICT_FP -c $S/rhs_builtin.cu -o $B/bdf_strict.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libbacon_ivp_$NAME.so $B/engine.o $S/build/rtc.o $B/fast.o $B/strict.o $B/bdf_fast.o $B/bdf_strict.o -ldl
grep -A2 "RkFastStepperINS_9RhsLorenzENS_8TabRKF45EEELb0" $B/ptxas_fast.log | grep -E "registers|spill"
