#!/usr/bin/env python
"""One table row per profiled launch of an `ncu --set full` report (time, grid, CTAs/SM, DRAM bytes,
L2->SM rate, tensor-pipe / issue utilisation, instructions), for profiles/*.txt.
usage: python tools/ncu_step_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
ci = {n: i for i, n in enumerate(h)}
cols = [("Kernel Name", "kernel", 26), ("gpu__time_duration.sum", "us", 7), ("launch__grid_size", "grid", 5),
        ("launch__registers_per_thread", "regs", 5), ("launch__occupancy_limit_shared_mem", "cta/SM(smem)", 12),
        ("dram__bytes_read.sum", "dram_rd", 9), ("dram__bytes_write.sum", "dram_wr", 9),
        ("derived__lts__lts2xbar_bytes.sum.per_second", "L2->SM TB/s", 11),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 8),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 7),
        ("smsp__inst_executed.sum", "warp-insts", 10)]
cols = [c for c in cols if c[0] in ci]
print("  ".join(f"{c[1]:>{c[2]}}" for c in cols))
units = rows[1]
for r in rows[2:]:
    out = []
    for name, _, w in cols:
        v = r[ci[name]]
        if name == "Kernel Name":
            v = v.replace("<unnamed>::", "").replace("void ", "").split("(")[0][:w]
        elif name.startswith("dram__"):
            v = f"{float(v):.2f}{units[ci[name]][0]}"
        else:
            try:
                v = f"{float(v):.2f}" if "." in v else v
            except ValueError:
                pass
        out.append(f"{v:>{w}}")
    print("  ".join(out))
