"""Sample error of the small sampler cases against the golden fixtures (max abs on x0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import cases
from tests.util import build_shell, golden

dev = torch.device("cuda:0")
for name in ("u_small_T20", "c_small_T20"):
    cfg = cases.SAMPLER_CASES[name]
    net, sd = build_shell(cfg, dev)
    x_T, noise, labels = cases.sampler_inputs(cfg)
    if cfg["kind"] == "uncond":
        from its_b200.Diffusion import GaussianDiffusionSampler
        smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"]).to(dev)
        x0 = smp(x_T.to(dev), noise=noise.to(dev))
    else:
        from its_b200.DiffusionFreeGuidence import GaussianDiffusionSampler
        smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"], w=cfg["w"]).to(dev)
        smp.print_steps = False
        x0 = smp(x_T.to(dev), labels.to(dev), noise=noise.to(dev))
    ref = torch.from_numpy(golden("smp_" + name)["x0"])
    d = (x0.cpu() - ref).abs()
    print(f"{name}: max abs {d.max().item():.5f}  mean abs {d.mean().item():.6f}  99.9% {d.flatten().kthvalue(int(d.numel()*0.999)).values.item():.5f}")
