#!/usr/bin/env python
"""Aggregate ncu warp-stall samples of one kernel by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_substring [top_n]
(needs the object compiled with -lineinfo; uses nvdisasm -g for the SASS offset -> line map)"""
import csv, re, subprocess, sys, tempfile, os, glob

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
cubin = glob.glob(os.path.join(td, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
off2line, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        infn = kname in ln
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# several kernels may be concatenated: take the block whose 'Kernel Name' row matches
blocks, curb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curb = {"name": r[1], "rows": []}
        blocks.append(curb)
    elif curb is not None:
        curb["rows"].append(r)
blk = [b for b in blocks if kname in b["name"]][0]
h = blk["rows"][0]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in blk["rows"][1:] if len(r) == len(h)]
base = min(int(r[ci["Address"]], 16) for r in body)
agg, tot = {}, 0.0
for r in body:
    s = float(r[ci["# Samples"]] or 0)
    tot += s
    key = off2line.get(int(r[ci["Address"]], 16) - base, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += s
    a[1] += float(r[ci["Instructions Executed"]] or 0)
lines = {}
for (f, l) in agg:
    if f != "?" and f not in lines:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f))
        lines[f] = open(cand[0]).read().splitlines() if cand else []
print(f"kernel {blk['name'][:60]}  total samples {tot:.0f}")
for (f, l), (s, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = lines.get(f, [])
    code = text[l - 1].strip()[:95] if 0 < l <= len(text) else ""
    print(f"{s / tot * 100:5.1f}%  exec={ex:>10.0f}  {f}:{l:<4d} {code}")
