"""What the two UNet shells share: plan cache, forward through the kernel plan.

The shells are torch.nn.Modules only so that parameters live under the
reference's state-dict keys (load_state_dict of a reference checkpoint works
unchanged, `module.`-stripped or not).  Their sub-modules are parameter
containers: no torch op of theirs runs on the sampling path."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import nn

from .engine import UNetPlan
from .engine_f32 import UNetPlanF32


class ParamOnly(nn.Module):
    """A sub-block whose arithmetic lives in the fused plan of the owning UNet."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is evaluated inside UNet.forward's kernel plan; "
                           "call the UNet, not the sub-module")


class PlannedUNet(nn.Module):
    is_conditional = False

    def _init_plans(self):
        self._plans: Dict[Tuple, UNetPlan] = {}
        self.register_load_state_dict_post_hook(lambda module, keys: module.invalidate_plans())

    def invalidate_plans(self) -> None:
        """Drop packed weights / buffers (call after mutating parameters in place)."""
        self._plans = {}

    def _apply(self, fn, *a, **k):  # .to() / .cuda() move the parameters: repack lazily
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def plan(self, n_img: int, H: int, W: int, *, n_img_in: Optional[int] = None, uniform_t: bool = False,
             impl: Optional[int] = None) -> UNetPlan:
        # `precision`: "16bit" (default; fp16 / bf16 operands on the tensor cores, fp32 accumulation: the
        # throughput path, samples within 2e-2 of the fp32 reference) or "fp32" (CUDA-core fp32 kernels, the parity
        # path: samples within 1e-4).  Set it on the model (`net.precision = "fp32"`) or through ITS_PRECISION.
        import os
        precision = str(getattr(self, "precision", None) or os.environ.get("ITS_PRECISION", "16bit")).lower()
        if precision not in ("16bit", "fp32"):
            raise ValueError(f"precision {precision!r}: '16bit' or 'fp32'")
        key = (n_img, H, W, n_img_in or n_img, uniform_t, impl, self.head.weight.data_ptr(),
               bool(getattr(self, "residual_fp16", False)), precision)
        p = self._plans.get(key)
        if p is None:
            cls = UNetPlanF32 if precision == "fp32" else UNetPlan
            p = cls(self, n_img, H, W, n_img_in=n_img_in, uniform_t=uniform_t, impl=impl)
            self._plans[key] = p
        return p

    def check_indices(self, t: Optional[torch.Tensor], labels: Optional[torch.Tensor]) -> None:
        """nn.Embedding raises IndexError on an index outside its table (the conditional net's time table
        `[T, ch]`, ModelCondition.py:38, and label table `[num_labels + 1, ch]`, :52-54): same error here, instead
        of reading a neighbouring row.  One host synchronisation per call (the samplers call it once per
        trajectory, not per step)."""
        if not self.is_conditional:
            return
        te = self.time_embedding.timembedding[0]
        if t is not None and t.numel() and bool(((t < 0) | (t >= te.num_embeddings)).any()):
            raise IndexError("index out of range in self")
        ce = self.cond_embedding.condEmbedding[0]
        if labels is not None and labels.numel() and bool(((labels < 0) | (labels >= ce.num_embeddings)).any()):
            raise IndexError("index out of range in self")

    def _run(self, x: torch.Tensor, t: torch.Tensor, labels: Optional[torch.Tensor]) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("its_b200 UNet runs on CUDA only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B,3,H,W], got {tuple(x.shape)}")
        if self.head.weight.device != x.device:
            raise RuntimeError(f"x lives on {x.device} but the UNet on {self.head.weight.device}")
        B, _, H, W = x.shape
        self.check_indices(t, labels)
        with torch.cuda.device(x.device):     # launches go to this device's current stream
            p = self.plan(B, H, W, impl=getattr(self, "impl", None))
            with torch.no_grad():
                p.x_in.copy_(x)
                p.t_idx.copy_(t.reshape(-1).to(torch.int64))
                if labels is not None:
                    p.labels.copy_(labels.reshape(-1).to(torch.int64))
            p.run_label_ops()
            p.run()
            return p.eps.clone()
