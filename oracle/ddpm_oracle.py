"""CPU oracle for the inference-time-scaling sampling path.  TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (torch CPU ops on plain state-dict tensors, no
nn.Module) of the reference's algorithm for this path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product package its_b200 never does.

Parity status: PINNED.  tests/golden/make_golden.py imports the reference's own
modules from /root/reference in the build container, runs them and this oracle
on identical weights/inputs, and commits the reference outputs as fixtures
under tests/golden/; tests/test_oracle_vs_golden.py (CPU) re-checks this file
against those fixtures.  The reference has no tests or golden vectors of its
own, so the fixtures are outputs of the reference itself.

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------- schedule --
def schedule(beta_1: float, beta_T: float, T: int) -> Dict[str, Tensor]:
    """fp64 schedule buffers.  Diffusion/Diffusion.py:57-65 (same in
    DiffusionFreeGuidence/DiffusionCondition.py:67-73): betas are an fp32
    linspace upcast to double; 'var' is the fixed-large variance of
    Diffusion.py:76."""
    betas = torch.linspace(beta_1, beta_T, T).double()
    alphas = 1.0 - betas
    alphas_bar = torch.cumprod(alphas, dim=0)
    alphas_bar_prev = torch.cat([torch.ones(1, dtype=torch.float64), alphas_bar])[:T]
    coeff1 = torch.sqrt(1.0 / alphas)
    coeff2 = coeff1 * (1.0 - alphas) / torch.sqrt(1.0 - alphas_bar)
    posterior_var = betas * (1.0 - alphas_bar_prev) / (1.0 - alphas_bar)
    var = torch.cat([posterior_var[1:2], betas[1:]])
    return dict(betas=betas, coeff1=coeff1, coeff2=coeff2, posterior_var=posterior_var, var=var,
                alphas_bar=alphas_bar)


# -------------------------------------------------------------------- UNet --
def _bf16(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _Q:
    """Optional emulation of the kernels' storage rounding (bf16 activations and
    conv weights, fp32 accumulation) used to separate 'kernel bug' from 'bf16'."""

    def __init__(self, mode: Optional[str]):
        self.on = mode == "bf16"

    def act(self, x: Tensor) -> Tensor:
        return _bf16(x) if self.on else x

    def w(self, x: Tensor) -> Tensor:
        return _bf16(x) if self.on else x


def _swish(x: Tensor) -> Tensor:  # Model.py:10-12
    return x * torch.sigmoid(x)


def _gn(sd, key: str, x: Tensor) -> Tensor:  # nn.GroupNorm(32, C): Model.py:132,170,186,258
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def _conv(sd, key: str, x: Tensor, q: _Q, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, q.w(sd[key + ".weight"]), sd[key + ".bias"], stride=stride, padding=padding)


def _linear(sd, key: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def _attn(sd, pre: str, x: Tensor, q: _Q) -> Tensor:
    """AttnBlock.forward, Model.py:145-164 (identical in ModelCondition.py:98-117)."""
    B, Cc, H, W = x.shape
    h = q.act(_gn(sd, pre + ".group_norm", x))
    qq = q.act(_conv(sd, pre + ".proj_q", h, q, padding=0))
    kk = q.act(_conv(sd, pre + ".proj_k", h, q, padding=0))
    vv = _conv(sd, pre + ".proj_v", h, q, padding=0)
    qq = qq.permute(0, 2, 3, 1).reshape(B, H * W, Cc)
    kk = kk.reshape(B, Cc, H * W)
    w = torch.bmm(qq, kk) * (int(Cc) ** (-0.5))
    w = F.softmax(w, dim=-1)
    vv = vv.permute(0, 2, 3, 1).reshape(B, H * W, Cc)
    if q.on:
        # kernel path: P and V are bf16 operands; V's bias is added after P@V
        # (rows of P sum to 1), Model.py:158-159.
        bias_v = sd[pre + ".proj_v.bias"]
        h = torch.bmm(_bf16(w), _bf16(vv - bias_v)) + bias_v
        h = _bf16(h)
    else:
        h = torch.bmm(w, vv)
    h = h.reshape(B, H, W, Cc).permute(0, 3, 1, 2)
    h = _conv(sd, pre + ".proj", h, q, padding=0)
    return q.act(x + h)


def _resblock(sd, pre: str, x: Tensor, temb: Tensor, cemb: Optional[Tensor], q: _Q) -> Tensor:
    """ResBlock.forward, Model.py:202-209; ModelCondition.py:153-161 adds cond_proj."""
    h = q.act(_swish(_gn(sd, pre + ".block1.0", x)))
    h = _conv(sd, pre + ".block1.2", h, q)
    h = h + _linear(sd, pre + ".temb_proj.1", _swish(temb))[:, :, None, None]
    if cemb is not None:
        h = h + _linear(sd, pre + ".cond_proj.1", _swish(cemb))[:, :, None, None]
    h = q.act(h)
    h = q.act(_swish(_gn(sd, pre + ".block2.0", h)))
    h = _conv(sd, pre + ".block2.3", h, q)  # block2 = GN, Swish, Dropout (eval: identity), Conv
    if (pre + ".shortcut.weight") in sd:
        h = h + _conv(sd, pre + ".shortcut", x, q, padding=0)
    else:
        h = h + x
    h = q.act(h)
    if (pre + ".attn.proj.weight") in sd:
        h = _attn(sd, pre + ".attn", h, q)
    return h


def time_embedding_uncond(sd, t: Tensor, ch: int) -> Tensor:
    """Functional sinusoid + MLP, Model.py:51-93."""
    emb = t.float().unsqueeze(-1) * sd["time_embedding.freq_coeffs"].unsqueeze(0)
    emb = torch.stack([torch.sin(emb), torch.cos(emb)], dim=-1).reshape(t.shape[0], ch)
    emb = _linear(sd, "time_embedding.timembedding.0", emb)
    return _linear(sd, "time_embedding.timembedding.2", _swish(emb))


def time_embedding_cond(sd, t: Tensor) -> Tensor:
    """Table embedding + MLP, ModelCondition.py:24-46."""
    emb = F.embedding(t, sd["time_embedding.timembedding.0.weight"])
    emb = _linear(sd, "time_embedding.timembedding.1", emb)
    return _linear(sd, "time_embedding.timembedding.3", _swish(emb))


def cond_embedding(sd, labels: Tensor) -> Tensor:
    """ConditionalEmbedding, ModelCondition.py:49-62 (padding_idx=0 only affects
    gradients; the forward is a plain lookup)."""
    emb = F.embedding(labels, sd["cond_embedding.condEmbedding.0.weight"])
    emb = _linear(sd, "cond_embedding.condEmbedding.1", emb)
    return _linear(sd, "cond_embedding.condEmbedding.3", _swish(emb))


def _count(sd, prefix: str) -> int:
    idx = {int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")}
    return max(idx) + 1 if idx else 0


def unet_forward(sd: Dict[str, Tensor], x: Tensor, t: Tensor, labels: Optional[Tensor] = None,
                 quant: Optional[str] = None) -> Tensor:
    """UNet.forward for both networks, driven by the state-dict keys alone.
    Model.py:265-285 (unconditional) / ModelCondition.py:206-235 (conditional,
    taken when `labels` is given)."""
    q = _Q(quant)
    cond = labels is not None
    ch = sd["head.weight"].shape[0]
    if cond:
        temb = time_embedding_cond(sd, t)
        cemb = cond_embedding(sd, labels)
    else:
        temb = time_embedding_uncond(sd, t, ch)
        cemb = None
    h = q.act(F.conv2d(x, sd["head.weight"], sd["head.bias"], padding=1))
    hs = [h]
    for i in range(_count(sd, "downblocks")):
        pre = f"downblocks.{i}"
        if (pre + ".block1.0.weight") in sd:
            h = _resblock(sd, pre, h, temb, cemb, q)
        elif cond:  # DownSample, ModelCondition.py:65-73
            h = q.act(_conv(sd, pre + ".c1", h, q, stride=2, padding=1) +
                      _conv(sd, pre + ".c2", h, q, stride=2, padding=2))
        else:       # DownSample, Model.py:96-108
            h = q.act(_conv(sd, pre + ".main", h, q, stride=2, padding=1))
        hs.append(h)
    for i in range(_count(sd, "middleblocks")):
        h = _resblock(sd, f"middleblocks.{i}", h, temb, cemb, q)
    for i in range(_count(sd, "upblocks")):
        pre = f"upblocks.{i}"
        if (pre + ".block1.0.weight") in sd:
            h = torch.cat([h, hs.pop()], dim=1)
            h = _resblock(sd, pre, h, temb, cemb, q)
        elif cond:  # UpSample, ModelCondition.py:76-86
            h = q.act(F.conv_transpose2d(h, q.w(sd[pre + ".t.weight"]), sd[pre + ".t.bias"], stride=2,
                                         padding=2, output_padding=1))
            h = q.act(_conv(sd, pre + ".c", h, q))
        else:       # UpSample, Model.py:111-126
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = q.act(_conv(sd, pre + ".main", h, q))
    assert len(hs) == 0
    h = q.act(_swish(_gn(sd, "tail.0", h)))
    return F.conv2d(h, sd["tail.2.weight"], sd["tail.2.bias"], padding=1)


# ------------------------------------------------------------------ sampler --
def p_mean_variance(sd, sched, x_t: Tensor, t: Tensor, labels: Optional[Tensor] = None, w: float = 0.0,
                    quant: Optional[str] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Diffusion.py:74-82 / DiffusionCondition.py:79-87.  Returns (mean, var, eps)."""
    var = sched["var"][t].float().view(-1, 1, 1, 1)
    if labels is None:
        eps = unet_forward(sd, x_t, t, quant=quant)
    else:
        e_c = unet_forward(sd, x_t, t, labels, quant=quant)
        e_u = unet_forward(sd, x_t, t, torch.zeros_like(labels), quant=quant)
        eps = (1.0 + w) * e_c - w * e_u
    c1 = sched["coeff1"][t].float().view(-1, 1, 1, 1)
    c2 = sched["coeff2"][t].float().view(-1, 1, 1, 1)
    return c1 * x_t - c2 * eps, var, eps


def sample(sd, sched, x_T: Tensor, noise_fn: Callable[[int], Tensor], labels: Optional[Tensor] = None,
           w: float = 0.0, quant: Optional[str] = None, t_start: Optional[int] = None,
           clip: bool = True, t_stop: int = 0) -> Tensor:
    """Ancestral loop with injected noise: Diffusion.py:84-102 /
    DiffusionCondition.py:89-105.  noise_fn(time_step) supplies z for steps
    T-1 .. 1 (the reference draws exactly T-1 tensors, none at step 0).
    t_start / t_stop cut the same loop into segments (steps t_start .. t_stop inclusive): the mid-trajectory
    restart of search over paths (BASELINE config 4) is two such segments with a perturbation in between."""
    T = sched["betas"].shape[0]
    x_t = x_T
    first = T - 1 if t_start is None else t_start
    for time_step in range(first, t_stop - 1, -1):
        t = torch.full((x_T.shape[0],), time_step, dtype=torch.long)
        mean, var, _ = p_mean_variance(sd, sched, x_t, t, labels, w, quant)
        if time_step > 0:
            x_t = mean + torch.sqrt(var) * noise_fn(time_step)
        else:
            x_t = mean
        assert torch.isnan(x_t).int().sum() == 0, "nan in tensor."
    return torch.clip(x_t, -1, 1) if clip else x_t


# ---------------------------------------------------------------- verifiers --
def oracle_verifier_score(images: Tensor) -> float:
    """OracleVerifier.score without dataset stats, search/verifier.py:60-63."""
    variance = torch.var(images.flatten(1), dim=1).mean().item()
    return 1.0 / (1.0 + variance)


def aesthetic_score(images: Tensor) -> float:
    """AestheticPredictor.score, search/verifier.py:277-287."""
    if images.min() < 0:
        images = (images + 1) / 2
    color_diversity = torch.std(images.flatten(1), dim=1).mean()
    contrast = torch.std(images.view(len(images), -1), dim=1).mean()
    return (color_diversity + contrast).item()


def self_supervised_score(images: Tensor) -> float:
    """SelfSupervisedVerifier.score with no reference features,
    search/verifier.py:218-221, 236-248."""
    f = F.adaptive_avg_pool2d(images, (8, 8)).flatten(1)
    f = F.normalize(f, dim=-1)
    sim = f @ f.T
    mask = ~torch.eye(len(f), dtype=torch.bool)
    return sim[mask].mean().item()


VERIFIERS = {"oracle": oracle_verifier_score, "aesthetic": aesthetic_score,
             "self_supervised": self_supervised_score}


# ------------------------------------------------------------------- search --
def random_search(noises: Sequence[Tensor], denoise_fn, verifier_fn) -> Tuple[int, float, List[float]]:
    """RandomSearch.search over pre-drawn candidates, search/search_algorithm.py:
    65-83: strict '>' keeps the first maximum."""
    best_i, best = -1, float("-inf")
    scores = []
    for i, z in enumerate(noises):
        s = verifier_fn(denoise_fn(z))
        scores.append(s)
        if s > best:
            best, best_i = s, i
    return best_i, best, scores


def zero_order_search(initial: Tensor, perturbations: Sequence[Sequence[Tensor]], lambda_radius: float,
                      denoise_fn, verifier_fn):
    """ZeroOrderSearch.search, search/search_algorithm.py:139-208, with the
    randn_like draws of _sample_neighbors (:223-229) supplied by the caller:
    perturbations[iteration][k] ~ N(0,1)."""
    current = initial.clone()
    best_noise, best = initial.clone(), float("-inf")
    history = {"scores": [], "candidates_per_iter": []}
    for pert in perturbations:
        neighbors = [current + p * (1 - lambda_radius) for p in pert]
        it_scores, it_best, it_best_noise = [], float("-inf"), None
        for nb in neighbors:
            s = verifier_fn(denoise_fn(nb))
            it_scores.append(s)
            if s > it_best:
                it_best, it_best_noise = s, nb.clone()
        history["scores"].append(it_scores)
        history["candidates_per_iter"].append(len(neighbors))
        if it_best > best:
            best, best_noise, current = it_best, it_best_noise.clone(), it_best_noise.clone()
    return best_noise, best, history


def path_search(initial: Tensor, variations: Sequence[Tensor], noise_scale: float, injection_step: int,
                denoise_fn, verifier_fn):
    """PathSearch.search, search/search_algorithm.py:296-336 (the reference's
    placeholder: perturb x_T, denoise fully, keep the best)."""
    best_noise, best = initial.clone(), float("-inf")
    history = {"scores": [], "injection_points": []}
    for v in variations:
        z = initial + v * noise_scale
        s = verifier_fn(denoise_fn(z))
        history["scores"].append(s)
        history["injection_points"].append(injection_step)
        if s > best:
            best, best_noise = s, z.clone()
    return best_noise, best, history


def path_search_restart(sd, sched, initial: Tensor, variations: Sequence[Tensor], noise_scale: float,
                        injection_step: int, noise_fn, verifier_fn, labels: Optional[Tensor] = None,
                        w: float = 0.0):
    """Search over paths with a real mid-trajectory restart (the paper's semantics, which the reference's
    placeholder at search/search_algorithm.py:307-312 announces but does not implement; BASELINE config 4):
    the pivot trajectory runs steps T-1 .. injection_step once; every path perturbs that intermediate state
    by noise_scale * variation and finishes steps injection_step-1 .. 0; strict '>' keeps the first best.
    Returns (best perturbed state, best score, history)."""
    base = sample(sd, sched, initial, noise_fn, labels, w, t_stop=injection_step, clip=False)
    best_state, best = base.clone(), float("-inf")
    history = {"scores": [], "injection_points": []}
    for v in variations:
        x = base + v * noise_scale
        s = verifier_fn(sample(sd, sched, x, noise_fn, labels, w, t_start=injection_step - 1))
        history["scores"].append(s)
        history["injection_points"].append(injection_step)
        if s > best:
            best, best_state = s, x.clone()
    return best_state, best, history


# ------------------------------------------------------------ synthetic init --
def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, gain: float = 1.0) -> Dict[str, Tensor]:
    """Deterministic, torch-RNG-independent weights for parity tests: every
    tensor with >= 2 dims ~ N(0, gain^2 / fan_in), biases ~ N(0, 0.02^2), norm
    weights 1 + N(0, 0.1^2).  Gives eps = O(1) so that a broken UNet cannot hide
    behind the reference's zero-gain initialisers (SURVEY.md §7 hard parts).
    Keys are visited in sorted order; `freq_coeffs` keeps its analytic value and
    sinusoid tables are rebuilt analytically."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("freq_coeffs"):
            d = shp[0] * 2
            out[k] = torch.exp(-(torch.arange(0, d, 2).float() / d * math.log(10000)))
        elif k == "time_embedding.timembedding.0.weight" and len(shp) == 2 and "time_embedding.timembedding.3.weight" in shapes:
            T, d = shp  # conditional net's sinusoid table, ModelCondition.py:27-34
            emb = torch.arange(0, d, 2) / d * math.log(10000)
            emb = torch.exp(-emb)
            pos = torch.arange(T).float()
            emb = pos[:, None] * emb[None, :]
            out[k] = torch.stack([torch.sin(emb), torch.cos(emb)], dim=-1).view(T, d).contiguous()
        elif len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            if k.endswith(".t.weight"):  # ConvTranspose2d: [in, out, kh, kw]
                fan_in = shp[0] * shp[2] * shp[3] // 4
            std = gain / math.sqrt(max(fan_in, 1))
            if k.startswith("cond_embedding.condEmbedding.0"):
                std = 1.0
            out[k] = torch.from_numpy((rng.standard_normal(shp) * std).astype(np.float32))
        elif ".group_norm.weight" in k or k.endswith((".block1.0.weight", ".block2.0.weight", "tail.0.weight")):
            out[k] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shp)).astype(np.float32))
        else:
            out[k] = torch.from_numpy((0.02 * rng.standard_normal(shp)).astype(np.float32))
    return out


def noise_stack(seed: int, T: int, shape: Tuple[int, ...]) -> Tensor:
    """[T, *shape] injected noise (entry t used at time_step t; entry 0 unused)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal((T,) + tuple(shape)).astype(np.float32))
