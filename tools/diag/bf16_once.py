"""One profiled launch of every bf16 layer kernel at the cfg3 shapes (warm-up launches outside the profiler range).
ncu --profile-from-start off --set full ... python tools/diag/bf16_once.py [B]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vqvae_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = 256
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
L = S // 4
calls = []


def layer(Cin, H, W, Cout, k, stride, transposed, out_f32=False):
    x = (torch.randn((B, H, W, Cin), device=dev, generator=g)).to(torch.bfloat16)
    wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
    w = torch.randn(wshape, device=dev, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn((Cout,), device=dev, generator=g) * 0.1
    kind = ops.conv_kind(k, stride, transposed, Cout)
    pk = ops.pack_conv_weight_bf16(w, kind)
    calls.append(lambda: ops.conv2d_bf16(x, pk, b, B=B, Cin=Cin, H=H, W=W, Cout=Cout, kind=kind, relu=True, out_f32=out_f32))


layer(64, S // 2, S // 2, 128, 4, 2, False)
layer(128, L, L, 128, 3, 1, False)
layer(128, L, L, 64, 1, 1, False, True)
layer(64, L, L, 128, 3, 1, True)
layer(128, L, L, 64, 4, 2, True)
layer(64, S // 2, S // 2, 3, 4, 2, True, True)
r = torch.randn((B, L, L, 128), device=dev, generator=g).clamp_min(0).to(torch.bfloat16)
w1 = torch.randn((32, 128, 3, 3), device=dev, generator=g) / np.sqrt(1152)
w2 = torch.randn((128, 32, 1, 1), device=dev, generator=g) / np.sqrt(32)
p1, p2 = ops.pack_conv_weight_bf16(w1, _lib.CONV_K3), ops.pack_conv_weight_bf16(w2, _lib.RES_W2)
calls.append(lambda: ops.residual_layer_bf16(r, p1, p2, B=B, H=L, W=L, C=128, Cmid=32, relu_out=True))
x = torch.rand((B, 3, S, S), device=dev, generator=g) * 2 - 1
wi = torch.randn((64, 3, 4, 4), device=dev, generator=g) / 7
pi = ops.pack_conv_weight(wi, False)
bi = torch.zeros((64,), device=dev)
calls.append(lambda: ops.conv_in_bf16(x, pi, bi, B=B, H=S, W=S, Cout=64))
for c in calls:
    c()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for c in calls:
    c()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
