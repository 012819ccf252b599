import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from vqvae_b200 import ops, _lib
from vqvae_b200._lib import TF32
rng = np.random.RandomState(0)
B,H,W,C,Cm = 256,8,8,128,32
r = torch.from_numpy(np.maximum(rng.standard_normal((B,H,W,C)).astype(np.float32),0)).cuda()
p1 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((Cm,C,3,3))/30).astype(np.float32)).cuda(), False)
p2 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((C,Cm,1,1))/6).astype(np.float32)).cuda(), False)
names = {0:'entry',1:'setup done',2:'first tiles landed',3:'chunk1',4:'chunk2',5:'chunk3',8:'gemm1 issued',9:'gemm1 done',10:'A2 written',11:'gemm2 issued',12:'gemm2 done',15:'act rewritten',25:'app1 gemm1 done',26:'app1 A2 written',28:'app1 gemm2 done',13:'epi2 stores issued',14:'exit'}
import os
NL = int(os.environ.get('NL', '2'))
for it in range(3):
    y = ops.residual_stack(r, p1, p2, B=B,H=H,W=W,C=C,Cmid=Cm, n_layers=NL, precision=TF32)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong*32)()
    _lib.check(_lib.lib().vqb_debug_read_trace(buf, 32), 'trace')
    t0 = buf[0]
    print('run', it, ' '.join(f"{names[i]}={ (buf[i]-t0)/1000:.2f}" for i in sorted(names, key=lambda i: buf[i]) if buf[i] >= t0))
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.residual_stack(r, p1, p2, B=B,H=H,W=W,C=C,Cmid=Cm, n_layers=NL, precision=TF32)
e1.record(); torch.cuda.synchronize(); print('avg per call (eager, incl launch gaps)', e0.elapsed_time(e1)/20*1000, 'us')
