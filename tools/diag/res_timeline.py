"""In-kernel timeline of res_bf16_kernel's CTA 0 at the cfg3 latent shape (diagnostic build: VQB_DIAG=1 python -m vqvae_b200.build)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vqvae_b200 import ops, _lib  # noqa: E402

assert ops.lib().vqb_diag_build() == 1, "build with VQB_DIAG=1"
B, L = 128, 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
r = torch.randn((B, L, L, 128), device=dev, generator=g).clamp_min(0).to(torch.bfloat16)
w1 = torch.randn((32, 128, 3, 3), device=dev, generator=g) / np.sqrt(1152)
w2 = torch.randn((128, 32, 1, 1), device=dev, generator=g) / np.sqrt(32)
p1, p2 = ops.pack_conv_weight_bf16(w1, _lib.CONV_K3), ops.pack_conv_weight_bf16(w2, _lib.RES_W2)
for _ in range(3):
    ops.residual_layer_bf16(r, p1, p2, B=B, H=L, W=L, C=128, Cmid=32, relu_out=True)
torch.cuda.synchronize()
fn = ops.lib().vqb_debug_read_res_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 512)()
assert fn(buf, 512) == 0
t = np.array(buf[:], dtype=np.int64).reshape(32, 16)
names = {0: "H halo c0 issue", 1: "S skip c0 issue", 2: "M g1 c0 start", 3: "M g1 c0 issued", 4: "M g1 c1 start", 5: "M g1 c1 issued",
         6: "M a2ready", 7: "M d2empty ok", 8: "E1 d1full seen", 9: "E1 done", 10: "E2 skip seen", 11: "E2 d2full seen",
         12: "E2 store issue", 13: "E2 store read"}
t0 = t[t > 0].min()
print("cycles relative to the first stamp; local tiles 5..12 of CTA 0")
print("%-16s" % "event" + "".join("%9d" % i for i in range(5, 13)))
for ev in sorted(names):
    print("%-16s" % names[ev] + "".join("%9d" % (t[i, ev] - t0 if t[i, ev] else -1) for i in range(5, 13)))
print("tile period (E2 store issue deltas):", np.diff(t[4:16, 12]))
