"""Verifiers with the reference's surface (search/verifier.py): OracleVerifier
(:30-66), SelfSupervisedVerifier (:191-248), AestheticPredictor (:251-287), each
`.score(images, ...) -> float`.  Scores are computed by the warp-shuffle
reduction kernels of libits_b200 (its_image_stats / its_candidate_scores); the
extra `.score_candidates(images, per_cand)` returns one score per candidate as a
device tensor so a whole population is scored without a host round-trip.

SupervisedVerifier, CLIPScore and IntegratedVerifier (:69-188, :290-388) wrap
openai-CLIP, which is not installed and whose weights cannot be fetched: parity
unpinned, kept importable, raise at construction when `clip` is missing.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch

from .. import _lib

KIND_ORACLE, KIND_AESTHETIC, KIND_SELFSUP = 0, 1, 2


def _stats(images: torch.Tensor, want_feats: bool):
    _lib.require_cuda()
    if images.device.type != "cuda":
        raise RuntimeError("its_b200 verifiers run on CUDA only (no CPU fallback)")
    if images.dim() != 4:
        raise ValueError(f"expected images [B,C,H,W], got {tuple(images.shape)}")
    x = images.detach().to(torch.float32).contiguous()
    n, c, h, w = x.shape
    stats = torch.empty((n, 4), dtype=torch.float32, device=x.device)
    feats = torch.empty((n, c * 64), dtype=torch.float32, device=x.device) if want_feats else None
    L = _lib.lib()
    with torch.cuda.device(x.device):
        _lib.check(L.its_image_stats(stats.data_ptr(), feats.data_ptr() if want_feats else None, x.data_ptr(),
                                     n, c, h, w, _lib.stream_ptr(x.device)), "its_image_stats")
    return stats, feats


def candidate_scores(images: torch.Tensor, per_cand: int, kind: int) -> torch.Tensor:
    """One fp32 score per group of `per_cand` consecutive images (device tensor)."""
    n = images.shape[0]
    if per_cand <= 0 or n % per_cand:
        raise ValueError(f"{n} images do not split into candidates of {per_cand}")
    stats, feats = _stats(images, kind == KIND_SELFSUP)
    scores = torch.empty((n // per_cand,), dtype=torch.float32, device=images.device)
    L = _lib.lib()
    with torch.cuda.device(images.device):
        _lib.check(L.its_candidate_scores(scores.data_ptr(), stats.data_ptr(),
                                          feats.data_ptr() if feats is not None else None, n // per_cand,
                                          per_cand, images.shape[1] * 64, kind, _lib.stream_ptr(images.device)),
                   "its_candidate_scores")
    return scores


class _KernelVerifier:
    kind = KIND_ORACLE

    def score_candidates(self, images: torch.Tensor, per_cand: int) -> torch.Tensor:
        return candidate_scores(images, per_cand, self.kind)


class OracleVerifier(_KernelVerifier):
    """Variance heuristic 1/(1+mean_b var(img_b)) when no dataset statistics are
    given (verifier.py:60-63); with statistics the reference's stub returns
    mean(images) (:65-66)."""
    kind = KIND_ORACLE

    def __init__(self, dataset_stats: Optional[Dict[str, np.ndarray]] = None):
        self.dataset_stats = dataset_stats

    def score(self, images: torch.Tensor, labels: Optional[torch.Tensor] = None) -> float:
        if self.dataset_stats is None:
            return float(self.score_candidates(images, images.shape[0]).item())
        stats, _ = _stats(images, False)
        return float(stats[:, 0].mean().item())   # equal-sized images: mean of means == global mean


class SelfSupervisedVerifier(_KernelVerifier):
    """Mean off-diagonal cosine similarity of L2-normalised 8x8 average-pooled
    images (verifier.py:207-248)."""
    kind = KIND_SELFSUP

    def __init__(self, denoising_features: Optional[torch.Tensor] = None):
        self.denoising_features = denoising_features

    def extract_features(self, images: torch.Tensor) -> torch.Tensor:
        """Un-normalised pooled features [B, C*64] (verifier.py:218-221)."""
        stats, feats = _stats(images, True)
        return feats * stats[:, 3:4]

    def score(self, images: torch.Tensor, reference_features: Optional[torch.Tensor] = None) -> float:
        if reference_features is None:
            return float(self.score_candidates(images, images.shape[0]).item())
        _, feats = _stats(images, True)                      # already normalised
        ref = torch.nn.functional.normalize(reference_features.to(feats), dim=-1)
        return torch.sum(feats * ref, dim=-1).item()         # raises for B > 1, like the reference


class AestheticPredictor(_KernelVerifier):
    """Contrast heuristic: 2 * mean_b std(img_b), on (x+1)/2 when the batch
    minimum is negative (verifier.py:277-287)."""
    kind = KIND_AESTHETIC

    def __init__(self, device: str = 'cuda'):
        self.device = device
        self.model = None

    def score(self, images: torch.Tensor) -> float:
        return float(self.score_candidates(images, images.shape[0]).item())


def _need_clip():
    try:
        import clip  # noqa: F401
    except ImportError as e:
        raise RuntimeError("this verifier wraps openai-CLIP, which is not installed here and has no "
                           "pinned weights (parity unpinned; outside the kernel scope)") from e


class SupervisedVerifier:
    def __init__(self, model_name: str = 'ViT-B/32', device: str = 'cuda', **kwargs: Any):
        _need_clip()
        raise NotImplementedError("CLIP-backed SupervisedVerifier is outside the accelerated path")


class CLIPScore:
    def __init__(self, device: str = 'cuda', **kwargs: Any):
        _need_clip()
        raise NotImplementedError("CLIP-backed CLIPScore is outside the accelerated path")


class IntegratedVerifier:
    def __init__(self, device: str = 'cuda', weights: Dict[str, float] = None):
        _need_clip()
        raise NotImplementedError("CLIP-backed IntegratedVerifier is outside the accelerated path")
