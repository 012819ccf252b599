"""In-kernel timeline of vq2_kernel's CTA 0 (needs the diagnostic build: VQB_DIAG=1 python -m vqvae_b200.build).
usage: VQB_DIAG=1 python -m vqvae_b200.build && python tools/diag/vq2_timeline.py [K] [N]
Prints, per local tile, the SM-cycle stamps of every pipeline hand-over relative to the first stamp."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vqvae_b200 import ops  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
assert ops.lib().vqb_diag_build() == 1, "build with VQB_DIAG=1"
dev = torch.device("cuda")
rng = np.random.RandomState(0)
z = torch.from_numpy(rng.standard_normal((N, 64)).astype(np.float32)).to(dev)
E = torch.from_numpy(rng.standard_normal((K, 64)).astype(np.float32)).to(dev)
for _ in range(3):
    ops.vq_forward(z, E)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 1024)()
fn = ops.lib().vqb_debug_read_vq2_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf, 1024) == 0
t = np.array(buf[:], dtype=np.int64).reshape(32, 32)
names = {0: "P load issue", 1: "P Q_DONE seen", 2: "P store read", 3: "M Z_FULL", 4: "M T_EMPTY c0", 5: "M commit c0", 6: "M T_EMPTY c1",
         7: "M commit c1", 8: "S Z_FULL", 9: "S T_FULL c0", 10: "S done c0", 11: "S T_FULL c1", 12: "S done c1", 13: "S exchanged",
         14: "S C_EMPTY", 15: "S C_FULL arr", 16: "S' done c0", 17: "S' done c1", 20: "F C_FULL", 21: "F bar1", 22: "F pairs", 23: "F bar3",
         24: "F emit done", 25: "F' emit done"}
t0 = t[t > 0].min()
order = sorted(names)
print("cycles relative to the first stamp; local tiles 4..11 of CTA 0 (K=%d, N=%d)" % (K, N))
print("%-14s" % "event" + "".join("%9d" % i for i in range(4, 12)))
for ev in order:
    print("%-14s" % names[ev] + "".join("%9d" % (t[i, ev] - t0 if t[i, ev] else -1) for i in range(4, 12)))
print("tile period (P load issue deltas):", np.diff(t[2:14, 0]))
