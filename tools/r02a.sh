#!/bin/bash
# round 2, call a: how the launch scales down (strong-scaling shards), current build vs 7 CTAs/SM; ncu of configs 4 and 5
for n in 131072 262144 524288 1048576; do
  BENCH_ARGS="--n $n" tools/ab.sh r02a_base_$n
  BENCH_ARGS="--n $n" tools/ab.sh r02a_minb7_$n BACON_IVP_LIB=variants/libbacon_ivp_minb7.so
done
BENCH_ARGS="--n 131072" tools/ab.sh r02a_base_notail_131072 BACON_IVP_NO_TAIL=1
BENCH_ARGS="--n 131072" tools/ab.sh r02a_minb7_notail_131072 BACON_IVP_NO_TAIL=1 BACON_IVP_LIB=variants/libbacon_ivp_minb7.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk_warp_linear32 -s 1 -c 1 -o gpurun_out/prof_cfg4_r02a -f python bench_configs.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_ncu_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/prof_cfg5_r02a -f python bench_configs.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_ncu_cfg5.log 2>&1
ls -la gpurun_out/*.ncu-rep
