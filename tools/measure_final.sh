#!/bin/bash
# tools/measure_final.sh TAG — whole GPU suite, headline bench (+ reference arm), ncu launch list of the bench command
T=${1:-rXX}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - "$T" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json")); r = d["roofline"]
print(f"bench: {d['value']:.4e} steps/s frac {r['frac']:.4f} kernel {r['kernel_ms_per_launch']:.3f} ms e2e {d['e2e']['value']:.4e} ({d['e2e']['ms_per_step']:.2f} ms) cpu {d['cpu_baseline']['value']:.3e} x{d['cpu_baseline']['cores']} launches {d['gpu_launches']} clocks {d['clocks']}")
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; cut -c1-260 gpurun_out/${T}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# one full ncu capture of the headline kernel (read here with profiles/summarize_ncu.py)
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 3 -c 1 -o gpurun_out/prof_rk45_${T} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu2.log 2>&1
