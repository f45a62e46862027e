#!/bin/bash
# tools/measure_paths.sh TAG — path-query tests (incl. linear32), timing legs on configs 2 and 4
T=${1:-rXX}
(timeout 400 python -m pytest tests/test_gpu_paths.py -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_paths_tests.log; tail -12 gpurun_out/${T}_paths_tests.log
for c in 4 2; do
timeout 300 python bench_configs.py --config $c --paths --steps 3 --no-cpu-baseline > gpurun_out/${T}_cfg${c}paths.json 2> gpurun_out/${T}_cfg${c}paths.err
python - "$T" $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg{sys.argv[2]}paths.json")); q = d["path_queries"]
    print(f"cfg{sys.argv[2]} paths:", d["value"], "steps/s;", {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3), v["regs_per_thread"]) for k, v in q.items() if isinstance(v, dict)}, q["events"]["events_found"])
except Exception as e:
    print("paths FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_cfg{sys.argv[2]}paths.err").read()[-2000:])
PY
done
# (the A/B runs of profiles/r01o_path_queries.md: tools/build_variant.sh NAME "-DBACON_EV_MINB=.. -DBACON_EV_UNROLL=.. -DBACON_EV_WIDE_MINB=..
#  -DBACON_EV_WIDE_STAGE=.." then BACON_IVP_LIB=variants/libbacon_ivp_NAME.so python bench_configs.py --config 2|4 --paths;
#  ncu: ncu --set full --clock-control none --import-source on -k regex:path_ -c 4 -o gpurun_out/prof_paths_TAG python bench_configs.py --config 2 --paths --steps 1 --no-cpu-baseline)
