#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + the hottest SASS lines by stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0       # which profiled launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + kidx]
print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.2:
        print(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# one block per profiled launch: a "Kernel Name" row, a header row, then the SASS lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
blk = blocks[kidx] if blocks else rows[1:]
h = blk[0]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in blk[1:] if len(r) == len(h)]
tot = sum(float(r[ci["# Samples"]] or 0) for r in body) or 1
print(f"\n-- top {top} SASS lines by samples (total {tot:.0f}) --")
for r in sorted(body, key=lambda r: -float(r[ci["# Samples"]] or 0))[:top]:
    print(f"{float(r[ci['# Samples']]) / tot * 100:5.1f}%  exec={r[ci['Instructions Executed']]:>9}  "
          f"thr={r[ci['Avg. Threads Executed']]:>4}  {r[ci['Source']].strip()[:90]}")
