# VQ-only timing probe: tc vs exact kernel, N = 2^20 rows
import sys, numpy as np, torch
sys.path.insert(0, '.')
from vqvae_b200 import ops
for K in (512, 1024):
    rng = np.random.RandomState(0)
    N = 1 << 20
    z = torch.from_numpy(rng.standard_normal((N, 64)).astype(np.float32)).cuda()
    E = torch.from_numpy(rng.standard_normal((K, 64)).astype(np.float32)).cuda()
    for kern in ("exact", "tc"):
        ops.set_vq_kernel(kern)
        for _ in range(3): ops.vq_forward(z, E)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.vq_forward(z, E)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = N * 520 / ms / 1e6
        print(f"K={K} {kern}: {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  {2*N*K*64/ms/1e9:.1f} TFLOP/s", flush=True)
ops.set_vq_kernel("auto")
