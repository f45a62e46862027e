#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize_ncu.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.md
  python profiles/summarize_ncu.py full gpurun_out/prof_rk45_r01.ncu-rep > profiles/r01_rk45_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        name = r["Kernel Name"]
        short = name.split("(")[0][-110:]
        a = agg.setdefault(short, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list ({path}): gpu__time_duration.sum, --clock-control none\n")
    print(f"{len(rows)} launches, {total:.3f} ms total device time (serialised, cold cache)\n")
    print("| kernel | launches | total ms | avg ms | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / a[0]:.4f} | {100 * a[1] / total:.2f}% | {a[2]} | {a[3]} |")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        print(f"# ncu --set full: {d.get('Kernel Name', ('?',))[0][:150]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d and d[k][0] != "":
                print(f"| {k} | {d[k][0]} | {d[k][1]} |")
        st = {h: float(v[0]) for h, v in d.items() if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v[0]}
        tot = sum(st.values()) or 1.0
        print("\n| warp stall reason (pc sampling) | samples | share |\n|---|---:|---:|")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
            print(f"| {k.replace('smsp__pcsamp_warps_issue_stalled_', '')} | {v:.0f} | {100 * v / tot:.1f}% |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
