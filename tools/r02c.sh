#!/bin/bash
# round 2, call c: what the regrouping kernel does at the end of an ensemble (ncu), the rest of the GPU tests, matvec layouts
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r02c_tests.log; cat gpurun_out/r02c_tests.log
tools/matvec_probe > gpurun_out/r02c_matvec_probe.txt 2>&1; cat gpurun_out/r02c_matvec_probe.txt
for n in 1048576 131072; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 3 -c 1 -o gpurun_out/prof_rk45_r02c_$n -f python bench.py --n $n --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_$n.log 2>&1
done
ls -la gpurun_out/*r02c*
