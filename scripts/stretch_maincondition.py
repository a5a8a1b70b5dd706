"""Stretch shape set: MainCondition.py's own defaults (ch_mult=[1,4,8,8,4,2], 547 M parameters, maps down to 1x1)
through the kernel plan against the CPU oracle.  Usage: python scripts/stretch_maincondition.py [B]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import ddpm_oracle as O
from its_b200.DiffusionFreeGuidence import UNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
MULT = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 8, 8, 4, 2]
IMG = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
cfg = dict(T=3000, num_labels=10, ch=128, ch_mult=MULT, num_res_blocks=2, dropout=0.15)
net = UNet(**cfg)
shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
sd = O.synth_state_dict(shapes, 5)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
print("params %.1f M" % (sum(p.numel() for p in net.parameters()) / 1e6))
g = np.random.default_rng(3)
x = torch.from_numpy(g.standard_normal((B, 3, IMG, IMG)).astype(np.float32))
t = torch.from_numpy(g.integers(0, cfg["T"], size=(B,)).astype(np.int64))
lab = torch.from_numpy(g.integers(0, 11, size=(B,)).astype(np.int64))
t0 = time.perf_counter()
with torch.no_grad():
    ref = O.unet_forward(sd, x, t, lab) if hasattr(O, "unet_forward") else None
print("oracle %.1f s" % (time.perf_counter() - t0))
with torch.no_grad():
    out = net(x.to(dev), t.to(dev), lab.to(dev)).cpu()
plan = next(iter(net._plans.values()))
kinds = {}
for k, _, n in plan.op_info:
    kinds[k] = kinds.get(k, 0) + n
print("launches", plan.n_launches, kinds)
if ref is not None:
    print("nan in out:", bool(torch.isnan(out).any()), "nan in oracle:", bool(torch.isnan(ref).any()), "max |oracle|", ref.abs().max().item())
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    print("max |out - oracle| / max |oracle| = %.3e" % err)
    assert err < 3e-2
print("ok")
