"""Two launches of the VQ kernel (N = 2^20 by default) for an ncu capture."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vqvae_b200 import ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
rng = np.random.RandomState(0)
z = torch.from_numpy(rng.standard_normal((N, 64)).astype(np.float32)).cuda()
E = torch.from_numpy(rng.standard_normal((K, 64)).astype(np.float32)).cuda()
ops.set_vq_kernel("tc")
for _ in range(2):
    ops.vq_forward(z, E)
torch.cuda.synchronize()
