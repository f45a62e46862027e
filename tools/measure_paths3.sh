#!/bin/bash
# tools/measure_paths3.sh TAG — path-query tests, A/B of the events kernel variants (tools/build_variant.sh), ncu captures
T=${1:-rXX}
(timeout 300 python -m pytest tests/test_gpu_paths.py -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_paths_tests.log; tail -5 gpurun_out/${T}_paths_tests.log
for v in default ev5 ev6 u2m8 u8m3; do
  L=""; [ $v != default ] && L=variants/libbacon_ivp_$v.so
  [ -n "$L" ] && [ ! -f "$L" ] && continue
  BACON_IVP_LIB=$L timeout 300 python bench_configs.py --config 2 --paths --steps 3 --no-cpu-baseline > gpurun_out/${T}_cfg2paths_$v.json 2> gpurun_out/${T}_cfg2paths_$v.err
  python - "$T" $v <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg2paths_{sys.argv[2]}.json")); q = d["path_queries"]
    print(f"{sys.argv[2]:>8}:", {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3), v["regs_per_thread"]) for k, v in q.items() if isinstance(v, dict)}, q["events"]["events_found"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_cfg2paths_{sys.argv[2]}.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:path_sample -c 1 -o gpurun_out/prof_sample_${T} -f python bench_configs.py --config 2 --paths --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_sample.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:path_events -c 1 -o gpurun_out/prof_events_${T} -f python bench_configs.py --config 2 --paths --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_events.log 2>&1
tail -2 gpurun_out/${T}_ncu_events.log
