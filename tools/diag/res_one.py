"""One residual_layer_bf16 launch at a given batch (cfg3 latent shape), for compute-sanitizer / ncu runs."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vqvae_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
r = torch.randn((B, L, L, 128), device=dev, generator=g).clamp_min(0).to(torch.bfloat16)
w1 = torch.randn((32, 128, 3, 3), device=dev, generator=g) / np.sqrt(1152)
w2 = torch.randn((128, 32, 1, 1), device=dev, generator=g) / np.sqrt(32)
p1, p2 = ops.pack_conv_weight_bf16(w1, _lib.CONV_K3), ops.pack_conv_weight_bf16(w2, _lib.RES_W2)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    y = ops.residual_layer_bf16(r, p1, p2, B=B, H=L, W=L, C=128, Cmid=32, relu_out=True)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
