"""Shared machinery of the two GaussianDiffusionSampler shells.

`forward` keeps the whole trajectory on the device: one CUDA graph holds
{UNet plan, fused DDPM step, step-counter decrement}; it is replayed T times with
no host synchronisation (the reference syncs on a NaN assert and prints every
step, Diffusion.py:91,100).  The NaN flag is OR-reduced on the device and checked
once per trajectory, raising the same AssertionError("nan in tensor.").
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from ._unet_base import PlannedUNet

X_T_TAG = 0x40000000  # Philox step tags >= this never collide with a time step


def extract(v, t, x_shape):
    """Coefficients at the given timesteps as [B,1,1,...] fp32 (Diffusion.py:9-16)."""
    out = torch.gather(v, index=t, dim=0).float().to(t.device)
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


class _Trajectory:
    """Per-(plan, noise-mode) device state + captured step graph."""

    def __init__(self):
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.noise_buf: Optional[torch.Tensor] = None
        self.key = None
        self.plan = None      # the captured graph bakes pointers into this plan: holding it keeps them alive


class SamplerBase(nn.Module):
    guided = False

    def _init_schedule(self, model, beta_1, beta_T, T):
        self.model = model
        self.T = T
        # fp32 linspace upcast to double, then fp64 throughout (Diffusion.py:57-65)
        self.register_buffer('betas', torch.linspace(beta_1, beta_T, T).double())
        alphas = 1. - self.betas
        alphas_bar = torch.cumprod(alphas, dim=0)
        alphas_bar_prev = F.pad(alphas_bar, [1, 0], value=1)[:T]
        self.register_buffer('coeff1', torch.sqrt(1. / alphas))
        self.register_buffer('coeff2', self.coeff1 * (1. - alphas) / torch.sqrt(1. - alphas_bar))
        self.register_buffer('posterior_var', self.betas * (1. - alphas_bar_prev) / (1. - alphas_bar))
        self.print_steps = True    # the reference prints every time step (Diffusion.py:91)
        self.use_cuda_graph = True
        self._coef_cache: Dict = {}
        self._traj: Dict = {}

    # ----------------------------------------------------------- public seam --
    def predict_xt_prev_mean_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return extract(self.coeff1, t, x_t.shape) * x_t - extract(self.coeff2, t, x_t.shape) * eps

    def _variance(self, x_t, t):
        var = torch.cat([self.posterior_var[1:2], self.betas[1:]])   # "fixed large", Diffusion.py:76
        return extract(var, t, x_t.shape)

    # -------------------------------------------------------------- internals --
    def _unet(self) -> PlannedUNet:
        m = self.model
        m = getattr(m, "module", m)            # tolerate a DataParallel-style wrapper
        if not isinstance(m, PlannedUNet):
            raise TypeError("its_b200 samplers drive an its_b200 UNet (kernel plan); got %r" % type(m).__name__)
        return m

    def _coef_table(self, dev) -> torch.Tensor:
        """[T][4] fp32 {c1, c2, sigma, 0}: the per-step scalars exactly as the
        reference forms them (fp64 gather -> .float(); sigma = sqrt of the fp32
        variance, Diffusion.py:13-15,76-77,99)."""
        key = (str(dev), self.T, self.betas.data_ptr())
        tab = self._coef_cache.get(key)
        if tab is None:
            var = torch.cat([self.posterior_var[1:2], self.betas[1:]]).float().to(dev)
            # sqrt on the target device, where the reference would evaluate torch.sqrt(var)
            tab = torch.stack([self.coeff1.float().to(dev), self.coeff2.float().to(dev), torch.sqrt(var),
                               torch.zeros_like(var)], dim=1).contiguous()
            self._coef_cache = {key: tab}
        return tab

    def _sample(self, x_T: torch.Tensor, labels: Optional[torch.Tensor], *, noise=None, seed=None,
                cand_id0: int = 0, t_start: Optional[int] = None, clip: bool = True,
                check_nan: bool = True, t_stop: int = 0) -> torch.Tensor:
        """Steps time_step = t_start (default T-1) ... t_stop (default 0) of the ancestral loop on the
        device; returns x_{t_stop - 1} (x_0, clipped when `clip`, if t_stop == 0).  Per-step noise is
        keyed by (seed, candidate, time_step), so a trajectory cut into segments equals the uncut one."""
        _lib.require_cuda()
        L = _lib.lib()
        if x_T.device.type != "cuda":
            raise RuntimeError("its_b200 samplers run on CUDA only (no CPU fallback)")
        with torch.cuda.device(x_T.device):   # every launch below goes to this device's current stream
            return self._sample_on_device(L, x_T, labels, noise=noise, seed=seed, cand_id0=cand_id0, t_start=t_start,
                                          clip=clip, check_nan=check_nan, t_stop=t_stop)

    def _sample_on_device(self, L, x_T, labels, *, noise, seed, cand_id0, t_start, clip, check_nan, t_stop):
        net = self._unet()
        if net.head.weight.device != x_T.device:
            raise RuntimeError(f"x_T lives on {x_T.device} but the UNet on {net.head.weight.device}")
        if self.guided:
            if getattr(net, "is_conditional", False) and self.T > net.time_embedding.timembedding[0].num_embeddings:
                raise IndexError("index out of range in self")     # the net's time table is shorter than the schedule
            net.check_indices(None, labels)
        B, Cc, H, W = x_T.shape
        n_net = 2 * B if self.guided else B
        plan = net.plan(n_net, H, W, n_img_in=B, uniform_t=True, impl=getattr(net, "impl", None))
        coef = self._coef_table(x_T.device)
        n_per = Cc * H * W
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # follows torch.manual_seed
        first = self.T - 1 if t_start is None else int(t_start)
        if not (0 <= first < self.T):
            raise ValueError(f"t_start={t_start} outside [0, {self.T})")
        t_stop = int(t_stop)
        if not (0 <= t_stop <= first):
            raise ValueError(f"t_stop={t_stop} outside [0, t_start={first}]")
        key = (noise is not None, clip, tuple(noise.shape) if noise is not None else None)
        tr = self._traj.get(key)
        if tr is not None and tr.plan is not plan:
            tr = None             # the plan was rebuilt (load_state_dict / .to() / new batch): never replay its graph
        if tr is None:
            tr = _Trajectory()
            tr.plan = plan
            tr.nan_flag = torch.zeros(1, dtype=torch.int32, device=x_T.device)
            tr.seed_args = None
            if noise is not None:
                tr.noise_buf = torch.empty_like(noise, dtype=torch.float32, device=x_T.device)
            self._traj = {key: tr}        # one live trajectory state per sampler
        with torch.no_grad():
            plan.x_in.copy_(x_T)
            plan.t_dev.fill_(first)
            tr.nan_flag.zero_()
            if self.guided:
                lab = labels.reshape(-1).to(device=x_T.device, dtype=torch.int64)
                plan.labels.copy_(torch.cat([lab, torch.zeros_like(lab)]))
                plan.run_label_ops()      # label embedding -> cond_proj vectors: once per trajectory
            if noise is not None:
                if noise.shape[0] != self.T or tuple(noise.shape[1:]) != tuple(x_T.shape):
                    raise ValueError("injected noise must be [T, *x_T.shape]; entry t is used at time_step t")
                tr.noise_buf.copy_(noise)
        eps_c = plan.eps.data_ptr()
        eps_u = plan.eps[B:].data_ptr() if self.guided else None
        w = float(getattr(self, "w", 0.0))
        noise_ptr = tr.noise_buf.data_ptr() if noise is not None else None
        stride = B * n_per if noise is not None else 0

        def step():
            plan.run()
            s = _lib.stream_ptr()
            _lib.check(L.its_ddpm_step(plan.x_in.data_ptr(), eps_c, eps_u, noise_ptr, stride, B, n_per,
                                       coef.data_ptr(), plan.t_dev.data_ptr(), w, seed, cand_id0,
                                       tr.nan_flag.data_ptr(), int(clip), s), "its_ddpm_step")
            _lib.check(L.its_step_advance(plan.t_dev.data_ptr(), -1, s), "its_step_advance")

        t_cur = first
        graph_ok = self.use_cuda_graph and first - t_stop >= 2
        # with injected noise the Philox key is never read: keep it out of the graph key
        graph_key = (None, 0, w) if noise is not None else (seed, cand_id0, w)
        if graph_ok and (tr.graph is None or tr.seed_args != graph_key):
            # (re)capture: seed / candidate base / guidance weight are baked into the graph.
            # One eager step first: it is a real step and it performs the lazy one-time
            # kernel attribute setup outside of stream capture.
            if self.print_steps:
                print(t_cur)
            step()
            t_cur -= 1
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            tr.graph, tr.seed_args = g, graph_key
        while t_cur >= t_stop:
            if self.print_steps:
                print(t_cur)
            if graph_ok:
                tr.graph.replay()
            else:
                step()
            t_cur -= 1
        if check_nan and int(tr.nan_flag.item()) != 0:
            raise AssertionError("nan in tensor.")
        self.last_launches_per_step = plan.n_launches + 2
        return plan.x_in.clone()
