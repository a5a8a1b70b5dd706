"""Class-conditional (classifier-free-guidance) UNet — same class names,
constructor signatures and parameter names as the reference's
DiffusionFreeGuidence/ModelCondition.py (UNet :164-235, ResBlock :121-161 with
cond_proj and attention on by default, DownSample = 3x3 s2 + 5x5 s2 :65-73,
UpSample = ConvTranspose2d(5,2,2,1) -> 3x3 :76-86, table TimeEmbedding :24-46,
ConditionalEmbedding with padding_idx=0 :49-62), default torch initialisers.

`UNet.forward(x, t, labels)` runs the sm_100a kernel plan; `return_representation`
(training analytics in the reference) is not part of the sampling path.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .._unet_base import ParamOnly, PlannedUNet


class Swish(ParamOnly):
    """x * sigmoid(x) (ModelCondition.py:19-21); fused into the kernels."""


class TimeEmbedding(ParamOnly):
    def __init__(self, T, d_model, dim):
        if d_model % 2:
            raise AssertionError("d_model must be even")
        super().__init__()
        freqs = torch.exp(-(torch.arange(0, d_model, step=2) / d_model * math.log(10000)))
        ang = torch.arange(T).float()[:, None] * freqs[None, :]
        table = torch.stack([torch.sin(ang), torch.cos(ang)], dim=-1).view(T, d_model)
        self.timembedding = nn.Sequential(nn.Embedding.from_pretrained(table, freeze=False),
                                          nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))


class ConditionalEmbedding(ParamOnly):
    def __init__(self, num_labels, d_model, dim):
        if d_model % 2:
            raise AssertionError("d_model must be even")
        super().__init__()
        self.condEmbedding = nn.Sequential(
            nn.Embedding(num_embeddings=num_labels + 1, embedding_dim=d_model, padding_idx=0),
            nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))


class DownSample(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        self.c1 = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1)
        self.c2 = nn.Conv2d(in_ch, in_ch, 5, stride=2, padding=2)


class UpSample(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        self.c = nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1)
        self.t = nn.ConvTranspose2d(in_ch, in_ch, 5, 2, 2, 1)


class AttnBlock(ParamOnly):
    def __init__(self, in_ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)


class ResBlock(ParamOnly):
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=True):
        super().__init__()
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(),
                                    nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.cond_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1))
        self.shortcut = (nn.Conv2d(in_ch, out_ch, 1, stride=1, padding=0) if in_ch != out_ch
                         else nn.Identity())
        self.attn = AttnBlock(out_ch) if attn else nn.Identity()


class UNet(PlannedUNet):
    is_conditional = True

    def __init__(self, T, num_labels, ch, ch_mult, num_res_blocks, dropout):
        super().__init__()
        tdim = ch * 4
        self.time_embedding = TimeEmbedding(T, ch, tdim)
        self.cond_embedding = ConditionalEmbedding(num_labels, ch, tdim)
        self.head = nn.Conv2d(3, ch, kernel_size=3, stride=1, padding=1)
        self.downblocks = nn.ModuleList()
        skip_chs, cur = [ch], ch
        for level, mult in enumerate(ch_mult):
            width = ch * mult
            for _ in range(num_res_blocks):
                self.downblocks.append(ResBlock(cur, width, tdim, dropout))  # attention at every down level
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
                self.upblocks.append(ResBlock(skip_chs.pop() + cur, width, tdim, dropout, attn=False))
                cur = width
            if level != 0:
                self.upblocks.append(UpSample(cur))
        assert len(skip_chs) == 0
        self.tail = nn.Sequential(nn.GroupNorm(32, cur), Swish(), nn.Conv2d(cur, 3, 3, stride=1, padding=1))
        self._init_plans()

    def forward(self, x, t, labels, return_representation=False):
        """eps = UNet(x_t, t, labels); labels [B] in 0..num_labels, 0 = null class."""
        if return_representation:
            raise NotImplementedError("return_representation is training analytics in the reference "
                                      "(ModelCondition.py:226-233) and outside the sampling path")
        return self._run(x, t, labels)
