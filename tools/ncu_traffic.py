"""Key an `ncu --set full` capture of tools/diag/step_once.py by bench.py's kernel labels.

  python tools/ncu_traffic.py gpurun_out/r02_step_kernels.ncu-rep gpurun_out/r02_step_labels.json \
      profiles/r02_step_kernels_traffic.json profiles/r02_step_kernels_ncu.txt

The report's launches are filtered to the layer kernels (one per label, in launch order); the JSON maps label -> DRAM
bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches with that label); the text file
is the per-launch table the design document cites."""
import csv
import json
import re
import subprocess
import sys

rep, labels_path, out_json, out_txt = sys.argv[1:5]
MAIN = re.compile(r"hconv_kernel|res_bf16_kernel|conv_in_bf16_kernel|vq2_kernel|vq_tc_kernel|vq_exact|conv_halo_kernel|conv_tc_kernel|"
                  r"res_tc_kernel|conv_in_tc_kernel|conv_ffma|conv_edge|res_ffma|convt_shuffle|conv_phase")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def scaled(r, name):
    """Metric value in base units (ncu prints bytes as Kbyte/Mbyte/..., times as usecond/msecond/...)."""
    if name not in col:
        return None
    v, u = r[col[name]].replace(",", ""), units[col[name]].lower()
    try:
        v = float(v)
    except ValueError:
        return None
    for pre, k in (("kbyte", 1e3), ("mbyte", 1e6), ("gbyte", 1e9), ("byte", 1.0), ("nsecond", 1e-3), ("usecond", 1.0), ("msecond", 1e3),
                   ("second", 1e6)):
        if u.startswith(pre):
            return v * k
    return v


launches = [r for r in rows[2:] if len(r) >= len(hdr) and MAIN.search(r[col["Kernel Name"]])]
labels = json.load(open(labels_path))
flat = [(cfg, lab) for cfg in ("cfg2", "cfg3") for lab in labels.get(cfg, [])]
if len(flat) != len(launches):
    sys.exit("label / launch count mismatch: %d labels, %d layer kernels: %s" % (len(flat), len(launches), [r[col["Kernel Name"]][:30] for r in launches]))
acc, lines = {}, []
lines.append("%-5s %-42s %-28s %9s %11s %11s %7s" % ("cfg", "bench label", "kernel", "us(ncu)", "dram rd MB", "dram wr MB", "issue%"))
for (cfg, lab), r in zip(flat, launches):
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    t = scaled(r, "gpu__time_duration.sum")
    acc.setdefault(lab, []).append(rd + wr)
    g = lambda n: (r[col[n]] if n in col else "-")  # noqa: E731
    kn = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    lines.append("%-5s %-42s %-28s %9.1f %11.2f %11.2f %7s" % (cfg, lab[:42], kn[:28], t, rd / 1e6, wr / 1e6,
                                                            g("smsp__issue_active.avg.pct_of_peak_sustained_active")[:6]))
json.dump({k: sum(v) / len(v) for k, v in acc.items()}, open(out_json, "w"), indent=1, sort_keys=True)
open(out_txt, "w").write("# ncu --set full --clock-control none, one VQVAE.forward per configuration (tools/diag/step_once.py); times are\n"
                         "# cold-cache and serialised -- shares, not absolutes, compare with bench.py's live CUDA-event times.\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
