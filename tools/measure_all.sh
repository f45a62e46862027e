#!/bin/bash
# tools/measure_all.sh TAG — on the GPU box: GPU tests, headline bench, launch list + full ncu capture of the headline
# kernel, and the other BASELINE configs at full size.  Everything lands in gpurun_out/ (summarised into profiles/ here).
T=${1:-rXX}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - "$T" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json")); r = d["roofline"]
print(f"bench: {d['value']:.4e} steps/s frac {r['frac']:.4f} kernel {r['kernel_ms_per_launch']:.3f} ms e2e {d['e2e']['value']:.4e} cpu {d['cpu_baseline']['value']:.3e} x{d['cpu_baseline']['cores']} clocks {d['clocks']}")
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 3 -c 1 -o gpurun_out/prof_rk45_${T} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu2.log 2>&1
for c in 3 4 5; do timeout 600 python bench_configs.py --config $c --steps 2 > gpurun_out/${T}_cfg$c.json 2> gpurun_out/${T}_cfg$c.err; python - "$T" $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg{sys.argv[2]}.json"))
    print(f"cfg{sys.argv[2]}: {d['value']:.4e} steps/s  {d['ms_per_pass']:.2f} ms  fp64 frac {d['roofline']['frac']:.3f}  failed {d['failed']}  dense {d.get('dense_output',{}).get('achieved_GBs')}  cpu {d.get('cpu_baseline',{}).get('value')}  parity {d.get('parity_sample')}")
except Exception as e:
    print("cfg", sys.argv[2], "FAILED", e)
PY
done
python bench_configs.py --config 5 --broyden --steps 2 --no-cpu-baseline > gpurun_out/${T}_cfg5_broyden.json 2>/dev/null
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline > gpurun_out/${T}_cfg2hist.json 2>/dev/null
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline --n 1048576 > gpurun_out/${T}_cfg2hist1m.json 2>/dev/null
python - "$T" <<'PY'
import json, sys
for c in ("cfg5_broyden", "cfg2hist", "cfg2hist1m"):
    try:
        d = json.load(open(f"gpurun_out/{sys.argv[1]}_{c}.json"))
        print(f"{c}: {d['value']:.4e} steps/s  {d['ms_per_pass']:.2f} ms  fp64 frac {d['roofline']['frac']:.3f}  dense {d.get('dense_output',{}).get('achieved_GBs')}")
    except Exception as e:
        print(c, "FAILED", e)
PY
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/prof_cfg2hist_${T} -f python bench_configs.py --config 2 --n 1048576 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_hist.log 2>&1
