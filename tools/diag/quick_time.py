# quick timing probe for the first GPU visit (not the bench)
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle.weights import make_images, make_state_dict
from tests.helpers import build_model
hp = dict(h_dim=128, res_h_dim=32, n_res_layers=2, n_embeddings=512, embedding_dim=64)
sd = make_state_dict(seed=0, **hp)
m = build_model(hp, sd)
for B,S in [(256,32),(32,256)]:
    x = torch.from_numpy(make_images(B,S,1)).cuda()
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m(x)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} S={S}: {e0.elapsed_time(e1)/10:.3f} ms/forward  {B/(e0.elapsed_time(e1)/10)*1e3:.0f} img/s", flush=True)
