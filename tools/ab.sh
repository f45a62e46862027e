#!/bin/bash
# tools/ab.sh LABEL [ENV=VAL ...] — one short headline bench run, prints value / roofline frac / kernel ms
L=$1; shift
env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/ab_$L.json 2> gpurun_out/ab_$L.err
python - "$L" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
    r = d["roofline"]
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} steps/s  frac {r['frac']:.4f}  kernel {r['kernel_ms_per_launch']:.3f} ms  e2e {d['e2e']['value']:.4e}  regs {d['arm']['regs_per_thread']} grid {d['arm']['grid']}x{d['arm']['block']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
