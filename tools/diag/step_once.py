"""One profiled VQVAE.forward per bench configuration (cfg2 in tf32 mode, cfg3 in bf16 mode), warm-ups outside the
profiler range; writes the per-kernel labels in launch order next to the report so tools/ncu_traffic.py can key the
DRAM traffic by bench.py's kernel labels.

  ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_step_kernels \
      python tools/diag/step_once.py gpurun_out/r02_step_labels.json
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (workload table + synthetic weights: the same model bench.py times)
import vqvae_b200  # noqa: E402
from vqvae_b200 import ops  # noqa: E402
from vqvae_b200.synth import make_images  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_step_labels.json"
dev = torch.device("cuda:0")
labels = {}
with torch.no_grad():
    for cfg, prec in (("cfg2", "tf32"), ("cfg3", "bf16")):
        wl = bench.WORKLOADS[cfg]
        model, _ = bench.build_model(wl, dev)
        x = torch.from_numpy(make_images(wl["batch"], wl["size"], seed=1)).to(dev)
        vqvae_b200.set_precision(prec)
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        ops.PROFILE = []
        torch.cuda.profiler.start()
        model(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        labels[cfg] = [lab for lab, _, _ in ops.PROFILE]
        ops.PROFILE = None
        del model, x
json.dump(labels, open(out_path, "w"), indent=1)
print({k: len(v) for k, v in labels.items()})
