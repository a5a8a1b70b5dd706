"""Full-length trajectory fixture: the REFERENCE sampler (Diffusion/Diffusion.py) on config A at its real width
(ch=128, ch_mult=[1,2,3,4], attn=[1]) for all T = 1000 steps, B = 2, synthetic O(1) weights and injected noise
rebuilt from numpy seeds.  Build container only (CPU, ~2-4 min):   python tests/golden/make_golden_long.py

The loop is driven through the public seam p_mean_variance exactly as Diffusion/Train.py:68-77 does (bit-identical
to sampler.forward), so that the un-clipped state can be stored as well.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import build_ref_model, ref_diffusion  # noqa: E402  (loads the reference modules by path)
from tests import cases  # noqa: E402

LONG_CASE = dict(cases.U_A, T=1000, beta_1=1e-4, beta_T=0.02, B=2, input_seed=601, noise_seed=602, weight_seed=61)


def main():
    torch.set_num_threads(8)
    cfg = LONG_CASE
    m, sd, _ = build_ref_model(cfg)
    smp = ref_diffusion.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"])
    x_T, noise, _ = cases.sampler_inputs(cfg)
    x_t = x_T
    keep = {}
    t0 = time.perf_counter()
    with torch.no_grad():
        for time_step in reversed(range(cfg["T"])):
            t = x_t.new_ones([x_T.shape[0]], dtype=torch.long) * time_step
            mean, var = smp.p_mean_variance(x_t=x_t, t=t)
            z = noise[time_step] if time_step > 0 else 0
            x_t = mean + torch.sqrt(var) * z
            if time_step in (900, 500, 100):
                keep[f"x_after_{time_step}"] = x_t.numpy().copy()
    print("reference trajectory: %.0f s" % (time.perf_counter() - t0))
    x0 = torch.clip(x_t, -1, 1)
    print("saturated fraction %.3f, pre-clip max %.3f" % ((x0.abs() == 1).float().mean(), x_t.abs().max()))
    np.savez_compressed(os.path.join(HERE, "smp_u_A_T1000.npz"), x0=x0.numpy(), x0_preclip=x_t.numpy(), **keep)


if __name__ == "__main__":
    main()
