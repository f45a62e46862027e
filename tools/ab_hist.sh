for v in cur h6 div2 h6div2; do
  if [ $v = cur ]; then L=; else L=variants/libbacon_ivp_$v.so; fi
  BACON_IVP_LIB=$L python bench_configs.py --config 2 --steps 2 --no-cpu-baseline --n 1048576 > gpurun_out/abh_$v.json 2>/dev/null
  python - $v <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/abh_{sys.argv[1]}.json")); print(sys.argv[1], '%.4e'%d['value'], round(d['ms_per_pass'],2), 'frac', round(d['roofline']['frac'],3), 'regs', d.get('regs_per_thread'), 'grid', d.get('grid'), 'dense', round((d.get('dense_output') or {}).get('achieved_GBs',0)))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tools/ab.sh head
