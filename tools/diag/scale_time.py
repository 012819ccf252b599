"""Diagnostic: kernel time vs number of CTAs (is weight streaming bound by aggregate L2 traffic or per SM?)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from vqvae_b200 import ops
from vqvae_b200._lib import TF32
rng = np.random.RandomState(0)
def t(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
C,Cm=128,32
w3 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((128,128,3,3))/30).astype(np.float32)).cuda(), False)
w4 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((128,64,4,4))/30).astype(np.float32)).cuda(), False)
p1 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((Cm,C,3,3))/30).astype(np.float32)).cuda(), False)
p2 = ops.pack_conv_weight(torch.from_numpy((rng.standard_normal((C,Cm,1,1))/6).astype(np.float32)).cuda(), False)
import inspect
print(inspect.signature(ops.conv2d))
for B in (16, 32, 64, 128, 256, 512):
    x8 = torch.from_numpy(np.maximum(rng.standard_normal((B,8,8,128)).astype(np.float32),0)).cuda()
    x16 = torch.from_numpy(rng.standard_normal((B,16,16,64)).astype(np.float32)).cuda()
    r = []
    r.append(t(lambda: ops.residual_stack(x8, p1, p2, B=B,H=8,W=8,C=C,Cmid=Cm, n_layers=2, precision=TF32)))
    try:
        r.append(t(lambda: ops.conv2d(x8, w3, None, B=B, Cin=128, H=8, W=8, Cout=128, kh=3, kw=3, stride=1, pad=1, transposed=False, relu=False, precision=TF32)))
        r.append(t(lambda: ops.conv2d(x16, w4, None, B=B, Cin=64, H=16, W=16, Cout=128, kh=4, kw=4, stride=2, pad=1, transposed=False, relu=True, precision=TF32)))
    except TypeError as e:
        print(e)
    print(f"B={B} tiles={B//2}: res_x2 {r[0]:.1f} us" + (f"  conv3x3 {r[1]:.1f} us  conv4x4s2 {r[2]:.1f} us" if len(r) > 1 else ""))
