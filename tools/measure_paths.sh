#!/bin/bash
# tools/measure_paths.sh TAG — on the GPU box: the path-query tests first, then the whole GPU suite, the path-query timing
# leg on config 2's history, and the headline bench.
T=${1:-rXX}
(timeout 300 python -m pytest tests/test_gpu_paths.py tests/test_gpu_plugin.py -m gpu -q 2>&1 | tail -40) > gpurun_out/${T}_paths_tests.log; tail -30 gpurun_out/${T}_paths_tests.log
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout 300 python bench_configs.py --config 2 --paths --steps 2 --no-cpu-baseline > gpurun_out/${T}_cfg2paths.json 2> gpurun_out/${T}_cfg2paths.err
python - "$T" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg2paths.json")); q = d["path_queries"]
    print("cfg2 paths:", d["value"], "steps/s;", {k: (round(v["ms"], 3), round(v["achieved_GBs"], 1), round(v["frac"], 3)) for k, v in q.items() if isinstance(v, dict)}, q["events"]["events_found"])
except Exception as e:
    print("cfg2paths FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_cfg2paths.err").read()[-2000:])
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - "$T" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json")); r = d["roofline"]
print(f"bench: {d['value']:.4e} steps/s frac {r['frac']:.4f} kernel {r['kernel_ms_per_launch']:.3f} ms e2e {d['e2e']['value']:.4e} ({d['e2e']['ms_per_step']:.2f} ms) cpu {d['cpu_baseline']['value']:.3e} x{d['cpu_baseline']['cores']} clocks {d['clocks']}")
PY
