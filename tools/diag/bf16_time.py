"""Per-layer device times of the bf16 kernels at a workload's shapes (CUDA events, 10 calls after 3 warm-ups)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vqvae_b200 import ops, _lib

B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def layer(name, Cin, H, W, Cout, k, stride, transposed, out_f32=False, relu=True):
    x = (torch.randn((B, H, W, Cin), device=dev, generator=g)).to(torch.bfloat16)
    wshape = (Cin, Cout, k, k) if transposed else (Cout, Cin, k, k)
    w = torch.randn(wshape, device=dev, generator=g) / np.sqrt(Cin * k * k)
    b = torch.randn((Cout,), device=dev, generator=g) * 0.1
    kind = ops.conv_kind(k, stride, transposed, Cout)
    pk = ops.pack_conv_weight_bf16(w, kind)
    us = t(lambda: ops.conv2d_bf16(x, pk, b, B=B, Cin=Cin, H=H, W=W, Cout=Cout, kind=kind, relu=relu, out_f32=out_f32))
    macs = B * H * W * Cin * Cout * k * k if transposed else B * (H // stride) * (W // stride) * Cin * Cout * k * k
    print(f"{name:28s} {us:9.1f} us   {2 * macs / us / 1e6:8.1f} TFLOP/s", flush=True)


L = S // 4
layer("E2 conv 64->128 k4s2", 64, S // 2, S // 2, 128, 4, 2, False)
layer("E3 conv 128->128 k3", 128, L, L, 128, 3, 1, False)
layer("pre 1x1 128->64 (f32 out)", 128, L, L, 64, 1, 1, False, True)
layer("D1 convT 64->128 k3", 64, L, L, 128, 3, 1, True)
layer("D2 convT 128->64 k4s2", 128, L, L, 64, 4, 2, True)
layer("D3 convT 64->3 k4s2 (gather form)", 64, S // 2, S // 2, 3, 4, 2, True, True)
layer("D3 convT 64->3 k4s2 (scatter form)", 64, S // 2, S // 2, 3, 4, 2, True, True, relu=False)
r = torch.randn((B, L, L, 128), device=dev, generator=g).clamp_min(0).to(torch.bfloat16)
w1 = torch.randn((32, 128, 3, 3), device=dev, generator=g) / np.sqrt(1152)
w2 = torch.randn((128, 32, 1, 1), device=dev, generator=g) / np.sqrt(32)
p1, p2 = ops.pack_conv_weight_bf16(w1, _lib.CONV_K3), ops.pack_conv_weight_bf16(w2, _lib.RES_W2)
us = t(lambda: ops.residual_layer_bf16(r, p1, p2, B=B, H=L, W=L, C=128, Cmid=32, relu_out=True))
fl = 2 * B * L * L * (9 * 128 * 32 + 32 * 128)
print(f"{'res 128->32->128':28s} {us:9.1f} us   {fl / us / 1e6:8.1f} TFLOP/s   {2 * B * L * L * 128 * 2 / us / 1e3:7.1f} GB/s", flush=True)

x = torch.rand((B, 3, S, S), device=dev, generator=g) * 2 - 1
wi = torch.randn((64, 3, 4, 4), device=dev, generator=g) / 7
pi = ops.pack_conv_weight(wi, False)
bi = torch.zeros((64,), device=dev)
us = t(lambda: ops.conv_in_bf16(x, pi, bi, B=B, H=S, W=S, Cout=64))
print(f"{'E1 conv 3->64 k4s2 (f32 in)':28s} {us:9.1f} us   {(B * 3 * S * S * 4 + B * S * S // 4 * 64 * 2) / us / 1e3:7.1f} GB/s", flush=True)
