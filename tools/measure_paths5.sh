#!/bin/bash
# tools/measure_paths5.sh TAG — linear32 query tests, then config 4's timing leg for the wide-record variants
T=${1:-rXX}
(timeout 400 python -m pytest tests/test_gpu_paths.py -m gpu -q -k linear32 2>&1 | tail -12) > gpurun_out/${T}_paths_tests.log; tail -4 gpurun_out/${T}_paths_tests.log
for v in default st6 d8 d4; do
  L=""; [ $v != default ] && L=variants/libbacon_ivp_$v.so
  [ -n "$L" ] && [ ! -f "$L" ] && continue
  BACON_IVP_LIB=$L timeout 300 python bench_configs.py --config 4 --paths --steps 3 --no-cpu-baseline > gpurun_out/${T}_cfg4paths_$v.json 2> gpurun_out/${T}_cfg4paths_$v.err
  python - "$T" $v <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg4paths_{sys.argv[2]}.json")); q = d["path_queries"]
    print(f"{sys.argv[2]:>8}:", {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3), v["regs_per_thread"]) for k, v in q.items() if isinstance(v, dict)}, q["events"]["events_found"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_cfg4paths_{sys.argv[2]}.err").read()[-1500:])
PY
done
