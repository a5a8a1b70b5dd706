"""Unconditional DDPM UNet — same class names, constructor signatures, parameter
names and initialisers as the reference's Diffusion/Model.py (UNet :212-285,
ResBlock :167-209, AttnBlock :129-164, Down/UpSample :96-126, TimeEmbedding
:15-93), so `load_state_dict(reference_checkpoint)` works unchanged.

The arithmetic does NOT live in these modules.  `UNet.forward(x, t)` runs a
launch plan of hand-written sm_100a kernels (its_b200.engine.UNetPlan):
tcgen05 implicit-GEMM convolutions fed by TMA, GroupNorm+Swish, attention,
embeddings.  The ViT of Model.py:289-456 is out of scope (never instantiated by
any driver of the reference).
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import init

from .._unet_base import ParamOnly, PlannedUNet


class Swish(ParamOnly):
    """x * sigmoid(x) (Model.py:10-12); fused into the GroupNorm / linear kernels."""


def _xavier(conv_or_linear, gain: float = 1.0):
    init.xavier_uniform_(conv_or_linear.weight, gain=gain)
    init.zeros_(conv_or_linear.bias)


class TimeEmbedding(ParamOnly):
    def __init__(self, T, d_model, dim):
        if d_model % 2:
            raise AssertionError("d_model must be even")
        super().__init__()
        self.d_model = d_model
        # functional sinusoid: frequencies only, valid for any t (Model.py:34-35)
        expo = torch.arange(0, d_model, step=2).float() / d_model * math.log(10000)
        self.register_buffer("freq_coeffs", torch.exp(-expo))
        self.timembedding = nn.Sequential(nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))
        for lin in (self.timembedding[0], self.timembedding[2]):
            _xavier(lin)


class DownSample(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        self.main = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1)
        _xavier(self.main)


class UpSample(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        self.main = nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1)
        _xavier(self.main)


class AttnBlock(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        # construction and initialisation order follow Model.py:130-144 draw for draw, so that
        # torch.manual_seed(s) + the constructor yields the reference's own parameters
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        for conv in (self.proj_q, self.proj_k, self.proj_v, self.proj):
            _xavier(conv)
        init.xavier_uniform_(self.proj.weight, gain=1e-5)


class ResBlock(ParamOnly):
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False):
        super().__init__()
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(),
                                    nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1))
        self.shortcut = (nn.Conv2d(in_ch, out_ch, 1, stride=1, padding=0) if in_ch != out_ch
                         else nn.Identity())
        self.attn = AttnBlock(out_ch) if attn else nn.Identity()
        # Model.py:199-204 re-initialises EVERY Conv2d / Linear below the block in module order — including
        # the attention block's projections, whose output projection therefore ends up with gain 1, not the
        # 1e-5 that AttnBlock.initialize had just given it — and then block2's conv with gain 1e-5.
        for module in self.modules():
            if isinstance(module, (nn.Conv2d, nn.Linear)):
                _xavier(module)
        init.xavier_uniform_(self.block2[-1].weight, gain=1e-5)


class UNet(PlannedUNet):
    is_conditional = False

    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout):
        super().__init__()
        assert all([i < len(ch_mult) for i in attn]), 'attn index out of bound'
        tdim = ch * 4
        self.time_embedding = TimeEmbedding(T, ch, tdim)
        self.head = nn.Conv2d(3, ch, kernel_size=3, stride=1, padding=1)
        self.downblocks = nn.ModuleList()
        skip_chs, cur = [ch], ch
        for level, mult in enumerate(ch_mult):
            width = ch * mult
            for _ in range(num_res_blocks):
                self.downblocks.append(ResBlock(cur, width, tdim, dropout, attn=(level in attn)))
                cur = width
                skip_chs.append(cur)
            if level != len(ch_mult) - 1:
                self.downblocks.append(DownSample(cur))
                skip_chs.append(cur)
        self.middleblocks = nn.ModuleList([ResBlock(cur, cur, tdim, dropout, attn=True),
                                           ResBlock(cur, cur, tdim, dropout, attn=False)])
        self.upblocks = nn.ModuleList()
        for level, mult in reversed(list(enumerate(ch_mult))):
            width = ch * mult
            for _ in range(num_res_blocks + 1):
                self.upblocks.append(ResBlock(skip_chs.pop() + cur, width, tdim, dropout, attn=(level in attn)))
                cur = width
            if level != 0:
                self.upblocks.append(UpSample(cur))
        assert len(skip_chs) == 0
        self.tail = nn.Sequential(nn.GroupNorm(32, cur), Swish(), nn.Conv2d(cur, 3, 3, stride=1, padding=1))
        _xavier(self.head)
        _xavier(self.tail[-1], gain=1e-5)
        self._init_plans()

    def forward(self, x, t):
        """eps = UNet(x_t, t); x [B,3,H,W] fp32 on CUDA, t [B] integer steps."""
        return self._run(x, t, None)
