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
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); ops.vq_forward(z, E); e1.record(); torch.cuda.synchronize(); print('one vq_forward call (3 launches + 1 memset):', e0.elapsed_time(e1)*1000, 'us')
buf = (ctypes.c_ulonglong * 48)()
_lib.check(_lib.lib().vqb_debug_read_trace_vq(buf, 48), 'trace')
k0 = buf[32]
print('kernel timeline CTA0: setup+norms %.1f us | tiles done %.1f us (%d tiles) | epilogue done %.1f | all warps done %.1f' % ((buf[33]-k0)/1000, (buf[34]-k0)/1000, buf[40], (buf[35]-k0)/1000, (buf[36]-k0)/1000))
names = ['tile start','z landed','wait T0','T0 ready','wait T1','T1 ready','pass1 done','filter+xchg done','zr loaded','rescored','best xchg done','emitted']
for t in range(2):
    b = buf[t*16]
    print('tile', t+1, ' | '.join(f"{names[i]} {(buf[t*16+i]-b)/1000:.2f}" for i in range(12)))
print('tile period us', (buf[16]-buf[0])/1000)

ct = (ctypes.c_ulonglong * 296)()
_lib.check(_lib.lib().vqb_debug_read_cta_times(ct, 296), 'cta')
st = np.array([ct[2*i] for i in range(148)], dtype=np.float64); en = np.array([ct[2*i+1] >> 10 for i in range(148)], dtype=np.float64)
smid = np.array([ct[2*i+1] & 1023 for i in range(148)])
t0 = st.min(); dur = (en - st) / 1000; print('CTA durations us: min %.0f med %.0f max %.0f' % (dur.min(), np.median(dur), dur.max()))
print('start offsets us: max %.1f' % ((st - t0).max() / 1000)); order = np.argsort(dur)
print('slowest CTAs (blockIdx, smid, dur):', [(int(i), int(smid[i]), round(float(dur[i]))) for i in order[-8:]])
print('fastest CTAs:', [(int(i), int(smid[i]), round(float(dur[i]))) for i in order[:8]])
print('distinct SMs', len(set(smid.tolist())))
