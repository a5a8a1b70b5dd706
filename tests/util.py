"""Helpers shared by the parity tests."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ddpm_oracle as O
from tests import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def build_shell(cfg, device=None):
    """its_b200 UNet shell with the case's synthetic weights (rebuilt from the seed)."""
    if cfg["kind"] == "uncond":
        from its_b200.Diffusion import UNet
        m = UNet(T=cfg["T"], ch=cfg["ch"], ch_mult=cfg["ch_mult"], attn=cfg["attn"],
                 num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"])
    else:
        from its_b200.DiffusionFreeGuidence import UNet
        m = UNet(T=cfg["T"], num_labels=cfg["num_labels"], ch=cfg["ch"], ch_mult=cfg["ch_mult"],
                 num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"])
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = O.synth_state_dict(shapes, cfg["weight_seed"])
    m.load_state_dict(sd, strict=True)
    m.eval()
    if device is not None:
        m = m.to(device)
    return m, sd


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|."""
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def rms_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a - b).pow(2).mean().sqrt()) / b.pow(2).mean().sqrt().clamp_min(1e-12))
