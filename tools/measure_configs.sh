T=r04s
for c in 3 4 5; do timeout 600 python bench_configs.py --config $c --steps 3 > gpurun_out/${T}_cfg$c.json 2> gpurun_out/${T}_cfg$c.err; python - "$T" $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_cfg{sys.argv[2]}.json"))
    print(f"cfg{sys.argv[2]}: {d['value']:.4e} steps/s  {d['ms_per_pass']:.2f} ms  fp64 frac {d['roofline']['frac']:.3f}  failed {d['failed']}  dense {d.get('dense_output',{}).get('achieved_GBs')}  cpu {d.get('cpu_baseline',{}).get('value')}  parity {d.get('parity_sample')}")
except Exception as e:
    print("cfg", sys.argv[2], "FAILED", e)
PY
done
bash tools/measure_paths.sh r04s 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:rk_warp_linear32 -c 1 -o gpurun_out/prof_cfg4_r04s -f python bench_configs.py --config 4 --steps 1 --no-cpu-baseline > gpurun_out/r04s_ncu_cfg4.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:path_ -c 6 --csv --log-file gpurun_out/r04s_paths_launches.csv python bench_configs.py --config 4 --paths --steps 1 --no-cpu-baseline > /dev/null 2>&1
