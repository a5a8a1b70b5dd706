"""The reference's config/*.yaml shape: Model.UNet(ch=128, ch_mult=[1,2,3,4], attn=[2]) on 3 x IMG x IMG images
(img_size: 256 -> 4096-token attention at level 2) through the kernel plan against the CPU oracle.
Usage: python scripts/stretch_img256.py [IMG] [B]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import ddpm_oracle as O
from its_b200.Diffusion import UNet

IMG = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
net = UNet(T=3000, ch=128, ch_mult=[1, 2, 3, 4], attn=[2], num_res_blocks=2, dropout=0.15)
shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
sd = O.synth_state_dict(shapes, 7)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
g = np.random.default_rng(4)
x = torch.from_numpy(g.standard_normal((B, 3, IMG, IMG)).astype(np.float32))
t = torch.from_numpy(g.integers(0, 3000, size=(B,)).astype(np.int64))
with torch.no_grad():
    out = net(x.to(dev), t.to(dev)).cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        net(x.to(dev), t.to(dev))
    torch.cuda.synchronize()
    print("gpu forward %.2f ms (eager launches, B=%d)" % ((time.perf_counter() - t0) / 3 * 1e3, B))
plan = next(iter(net._plans.values()))
kinds = {}
for k, _, n in plan.op_info:
    kinds[k] = kinds.get(k, 0) + n
print("launches", plan.n_launches, kinds)
t0 = time.perf_counter()
with torch.no_grad():
    ref = O.unet_forward(sd, x, t, None)
print("oracle %.1f s" % (time.perf_counter() - t0))
print("nan in out:", bool(torch.isnan(out).any()))
err = (out - ref).abs().max().item() / ref.abs().max().item()
print("max |out - oracle| / max |oracle| = %.3e" % err)
assert err < 3e-2
print("ok")
