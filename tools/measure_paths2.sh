#!/bin/bash
# tools/measure_paths2.sh TAG — path-query tests, timing leg, and one ncu --set full capture of both query kernels
T=${1:-rXX}
(timeout 300 python -m pytest tests/test_gpu_paths.py -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_paths_tests.log; tail -12 gpurun_out/${T}_paths_tests.log
timeout 300 python bench_configs.py --config 2 --paths --steps 3 --no-cpu-baseline > gpurun_out/${T}_cfg2paths.json 2> gpurun_out/${T}_cfg2paths.err
python - "$T" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg2paths.json")); q = d["path_queries"]
    print("cfg2 paths:", d["value"], "steps/s;", {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3), v["regs_per_thread"]) for k, v in q.items() if isinstance(v, dict)}, q["events"]["events_found"])
except Exception as e:
    print("cfg2paths FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_cfg2paths.err").read()[-2000:])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:path_ -c 4 -o gpurun_out/prof_paths_${T} -f python bench_configs.py --config 2 --paths --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_paths.log 2>&1
tail -3 gpurun_out/${T}_ncu_paths.log
