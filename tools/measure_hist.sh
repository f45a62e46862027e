python bench_configs.py --config 5 --steps 2 --no-cpu-baseline > gpurun_out/r01i_cfg5.json 2>gpurun_out/r01i_cfg5.err
python bench_configs.py --config 5 --broyden --steps 2 --no-cpu-baseline > gpurun_out/r01i_cfg5_broyden.json 2>/dev/null
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline > gpurun_out/r01i_cfg2hist.json 2>/dev/null
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline --no-hist > gpurun_out/r01i_cfg2nohist.json 2>/dev/null
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline --n 1048576 > gpurun_out/r01i_cfg2hist1m.json 2>gpurun_out/r01i_cfg2hist1m.err
python bench_configs.py --config 2 --steps 2 --no-cpu-baseline --n 1048576 --no-hist > gpurun_out/r01i_cfg2nohist1m.json 2>/dev/null
python bench_configs.py --config 4 --steps 2 --no-cpu-baseline --no-hist > gpurun_out/r01i_cfg4nohist.json 2>/dev/null
for f in cfg5 cfg5_broyden cfg2hist cfg2nohist cfg2hist1m cfg2nohist1m cfg4nohist; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/r01i_{sys.argv[1]}.json")); print(sys.argv[1], '%.4e'%d['value'], round(d['ms_per_pass'],2), 'frac', round(d['roofline']['frac'],3), 'regs', d.get('regs_per_thread'), 'grid', d.get('grid'), 'dense', (d.get('dense_output') or {}).get('achieved_GBs'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 1 -c 1 -o gpurun_out/prof_cfg2hist_r01i -f python bench_configs.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01i_ncu_hist.log 2>&1
