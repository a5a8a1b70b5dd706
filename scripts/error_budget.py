"""CPU study with the oracle: which 16-bit storage choice contributes how much to the error of one UNet
evaluation (config A, synthetic O(1) weights)?  Sites: GroupNorm(+Swish) outputs and the weights that multiply
them ("gn"), raw feature maps — conv1 outputs, the residual stream, q/k/v/P ("raw") — and the weights of the
1x1 shortcut / projection convs that multiply raw maps.  Usage: python scripts/error_budget.py"""
import linecache
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ddpm_oracle as O
from tests import cases
from tests.util import rel_err, rms_err

FMT = {"fp32": lambda x: x, "bf16": lambda x: x.to(torch.bfloat16).float(), "fp16": lambda x: x.to(torch.float16).float()}


BOUNDED_LINES = {86, 87, 114}      # q, k projections and conv1 outputs: inputs are GroupNorm outputs -> bounded


class Q(O._Q):
    """gn: GroupNorm(+Swish) outputs; raw: the residual stream (block / resample / attention outputs, head);
    mid: raw maps whose inputs are GroupNorm outputs (conv1 outputs, q, k) and are therefore bounded."""

    def __init__(self, gn, raw, w, mid=None):
        self.on = True
        self.gn, self.raw, self.wf, self.mid = FMT[gn], FMT[raw], FMT[w], FMT[mid or raw]

    def act(self, x):
        f = sys._getframe(1)
        line = linecache.getline(f.f_code.co_filename, f.f_lineno)
        if "_gn(" in line:
            return self.gn(x)
        return self.mid(x) if f.f_lineno in BOUNDED_LINES else self.raw(x)

    def w(self, x):
        return self.wf(x)


cfg = dict(cases.U_A, B=2)
torch.set_num_threads(8)
import numpy as np
from tests.util import build_shell
net, sd = build_shell(cfg, None)
x, t, labels = cases.forward_inputs(cfg)
orig_Q = O._Q
with torch.no_grad():
    ref = O.unet_forward(sd, x, t, labels)
    for gn, raw, w, mid in [("fp16", "bf16", "fp16", None), ("bf16", "bf16", "bf16", None), ("fp16", "fp16", "fp16", None),
                            ("fp16", "bf16", "fp16", "fp16"), ("fp16", "fp32", "fp16", "bf16"), ("fp32", "bf16", "fp32", "fp32"),
                            ("fp32", "fp32", "fp32", "bf16"), ("fp16", "fp32", "fp16", None)]:
        q = Q(gn, raw, w, mid)
        O._Q = lambda mode, q=q: q          # unet_forward builds its _Q from the `quant` argument
        out = O.unet_forward(sd, x, t, labels, quant="bf16")
        O._Q = orig_Q
        print(f"gn outputs {gn:5s} residual stream {raw:5s} bounded raw maps {mid or raw:5s} weights {w:5s}: rms rel {rms_err(out, ref):.2e}  max rel {rel_err(out, ref):.2e}")
