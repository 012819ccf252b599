import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from vqvae_b200 import ops
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.RandomState(0)
N = 1 << 20
z = torch.from_numpy(rng.standard_normal((N, 64)).astype(np.float32)).cuda()
E = torch.from_numpy(rng.standard_normal((K, 64)).astype(np.float32)).cuda()
ops.set_vq_kernel("tc")
for _ in range(3): ops.vq_forward(z, E)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.vq_forward(z, E)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"K={K} flags={os.environ.get('VQB_TC_FLAGS','0')}: {ms:.3f} ms  {N*520/ms/1e6:.0f} GB/s", flush=True)
