"""Diagnostic: VQ kernel time and slow-path sensitivity over several cfg2 batches (different image seeds)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import vqvae_b200
from vqvae_b200 import ops
from oracle import weights
sd = weights.make_state_dict(128, 32, 2, 512, 64, seed=0)
m = vqvae_b200.VQVAE(128, 32, 2, 512, 64, 0.25)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
vqvae_b200.set_precision("tf32")
m = m.cuda().eval()
cb = m.vector_quantization._codebook()
for seed in (1, 2, 3, 101, 102, 103, 108):
    x = torch.from_numpy(weights.make_images(256, 32, seed=seed)).cuda()
    with torch.no_grad():
        z_e, B, H, W = m._encode_rows(x)
    rows = z_e.view(-1, 64)
    for _ in range(3): ops.vq_forward(rows, cb)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.vq_forward(rows, cb)
    e1.record(); torch.cuda.synchronize()
    ops.set_vq_kernel("exact"); i0 = ops.vq_forward(rows, cb)[0]; ops.set_vq_kernel("auto")
    i1 = ops.vq_forward(rows, cb)[0]
    print(f"seed {seed}: vq_forward {e0.elapsed_time(e1)/20*1000:.1f} us per call, idx equal to exact kernel: {bool(torch.equal(i0, i1))}")
