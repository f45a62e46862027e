#!/bin/bash
# round 2, call b: in-kernel regrouping (one wide CTA per SM): GPU tests, then the launch scaled down
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02b_tests.log; cat gpurun_out/r02b_tests.log
for n in 131072 262144 524288 1048576; do
  BENCH_ARGS="--n $n" tools/ab.sh r02b_$n
done
BENCH_ARGS="--n 131072" tools/ab.sh r02b_nomig_131072 BACON_IVP_LIB=variants/libbacon_ivp_nomig.so
BENCH_ARGS="--n 1048576" tools/ab.sh r02b_nomig_1048576 BACON_IVP_LIB=variants/libbacon_ivp_nomig.so
