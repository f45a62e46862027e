T=r04b
(timeout 400 python -m pytest tests/test_gpu_paths.py tests/test_gpu_plugin.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/${T}_paths_tests.log; cat gpurun_out/${T}_paths_tests.log
run() { # name lib config
  BACON_IVP_LIB=$2 timeout 300 python bench_configs.py --config $3 --paths --steps 3 --no-cpu-baseline > gpurun_out/${T}_cfg$3paths_$1.json 2> gpurun_out/${T}_cfg$3paths_$1.err
  python - gpurun_out/${T}_cfg$3paths_$1.json $1 $3 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); q = d["path_queries"]
    print(sys.argv[2], "cfg", sys.argv[3], {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3)) for k, v in q.items() if isinstance(v, dict)})
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run i64 bacon_b200/libbacon_ivp.so 2
for v in bis i32 i128 i4096; do run $v variants/libbacon_ivp_$v.so 2; done
run i64 bacon_b200/libbacon_ivp.so 4
run bis variants/libbacon_ivp_bis.so 4
run i4096 variants/libbacon_ivp_i4096.so 4
