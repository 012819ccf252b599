import sys, os, ctypes, numpy as np, torch
os.environ["VQB_TC_FLAGS"] = "8"
sys.path.insert(0, '.')
from vqvae_b200 import ops, _lib
rng = np.random.RandomState(0)
N, K = 1 << 20, 512
z = torch.from_numpy(rng.standard_normal((N, 64)).astype(np.float32)).cuda()
E = torch.from_numpy(rng.standard_normal((K, 64)).astype(np.float32)).cuda()
ops.set_vq_kernel("tc")
for _ in range(3): ops.vq_forward(z, E)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)()
_lib.check(_lib.lib().vqb_debug_read_trace_vq(buf, 32), 'trace')
names = ['tile start','z landed','wait T0','T0 ready','wait T1','T1 ready','pass1 done','filter+xchg done','zr loaded','rescored','best xchg done','emitted']
for t in range(2):
    b = buf[t*16]
    print('tile', t+1, ' | '.join(f"{names[i]} {(buf[t*16+i]-b)/1000:.2f}" for i in range(12)))
print('tile period us', (buf[16]-buf[0])/1000)
