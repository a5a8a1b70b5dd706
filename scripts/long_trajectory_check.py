"""Full-length (T = 1000) config-A trajectory against the reference fixture tests/golden/smp_u_A_T1000.npz
(same synthetic weights, x_T and injected noise): error of the un-clipped state along the trajectory and of the
clipped samples."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import torch

from tests import cases
from tests.util import build_shell, golden
from its_b200.Diffusion import GaussianDiffusionSampler

LONG_CASE = dict(cases.U_A, T=1000, beta_1=1e-4, beta_T=0.02, B=2, input_seed=601, noise_seed=602, weight_seed=61)
dev = torch.device("cuda:0")
g = golden("smp_u_A_T1000")
net, _ = build_shell(LONG_CASE, dev)
smp = GaussianDiffusionSampler(net, 1e-4, 0.02, 1000).to(dev)
smp.print_steps = False
x_T, noise, _ = cases.sampler_inputs(LONG_CASE)
x_T, noise = x_T.to(dev), noise.to(dev)
x = x_T
first = 999
for stop in (900, 500, 100, 0):
    x = smp(x, noise=noise, t_start=first, t_stop=stop, clip=False)
    ref = torch.from_numpy(g["x0_preclip"] if stop == 0 else g[f"x_after_{stop}"]).to(dev)
    rel = ((x - ref).abs().max() / ref.abs().max()).item()
    rms = ((x - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"after step {stop:4d}: max|ref| {ref.abs().max().item():9.3f}  max-abs err / max|ref| {rel:.3e}  rms rel {rms:.3e}")
    first = stop - 1
x0 = torch.clip(x, -1, 1)
ref0 = torch.from_numpy(g["x0"]).to(dev)
d = (x0 - ref0).abs()
unsat = ref0.abs() < 1
print(f"clipped samples: max abs diff {d.max().item():.3e}; pixels over 2e-2: {(d > 2e-2).float().mean().item():.2e} "
      f"({int((d > 2e-2).sum())} of {d.numel()}); unsaturated reference pixels: {int(unsat.sum())}")
