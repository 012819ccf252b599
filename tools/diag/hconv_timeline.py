"""In-kernel timeline of hconv_kernel's CTA 0 for the cfg3 conv layers (diagnostic build: VQB_DIAG=1 python -m vqvae_b200.build).
Prints SM-cycle stamps per local tile: halo loads issued, weight groups issued, MMA wait / first / last, epilogue start / end."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vqvae_b200 import ops  # noqa: E402

assert ops.lib().vqb_diag_build() == 1, "build with VQB_DIAG=1"
B, S = 128, 256
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
L = S // 4
fn = ops.lib().vqb_debug_read_hconv_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = {0: "H chunk0 issue", 1: "H chunk1 issue", 2: "W first group", 3: "W last group", 4: "M tile start", 5: "M tempty ok",
         6: "M halo ok", 7: "M last issued", 8: "E tfull seen", 9: "E done (w4)", 10: "E done (w11)"}


def layer(name, Cin, H, W, Cout, k, stride, transposed, out_f32=False, relu=True):
    x = (torch.randn((B, H, W, Cin), device=dev, generator=g)).to(torch.bfloat16)
    wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
    w = torch.randn(wshape, device=dev, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn((Cout,), device=dev, generator=g) * 0.1
    kind = ops.conv_kind(k, stride, transposed, Cout)
    pk = ops.pack_conv_weight_bf16(w, kind)
    for _ in range(3):
        ops.conv2d_bf16(x, pk, b, B=B, Cin=Cin, H=H, W=W, Cout=Cout, kind=kind, relu=relu, out_f32=out_f32)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 512)()
    assert fn(buf, 512) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(32, 16)
    t0 = t[t > 0].min()
    print("== %s: cycles relative to the first stamp, local tiles 3..9 of CTA 0" % name)
    print("%-16s" % "event" + "".join("%9d" % i for i in range(3, 10)))
    for ev in sorted(names):
        print("%-16s" % names[ev] + "".join("%9d" % (t[i, ev] - t0 if t[i, ev] else -1) for i in range(3, 10)))
    print("tile period (M tile start deltas):", np.diff(t[2:12, 4]))


layer("E2 conv 64->128 k4s2", 64, S // 2, S // 2, 128, 4, 2, False)
layer("E3 conv 128->128 k3", 128, L, L, 128, 3, 1, False)
layer("D1 convT 64->128 k3", 64, L, L, 128, 3, 1, True)
layer("D2 convT 128->64 k4s2", 128, L, L, 64, 4, 2, True)
layer("pre 1x1 128->64", 128, L, L, 64, 1, 1, False, True, False)
