"""Search over noise candidates with the reference's surface
(search/search_algorithm.py): RandomSearch (:18-87), ZeroOrderSearch (:90-235),
PathSearch (:238-340); same constructors, `.search(...)` signatures, return
values, `.nfes` / `.reset_nfes()`.

Two execution modes
  * population mode — taken when `denoise_fn` is an its_b200 `SamplerDenoiser`
    (see `make_denoise_fn`) and `verifier_fn` is the bound `.score` of an
    its_b200 verifier: the whole candidate population of a round is denoised as
    device batches (candidates are the batch dimension), scored on the device,
    and selected with the first-index argmax kernel.  With torch.distributed
    initialised the candidates are block-sharded over the ranks; the only
    collective is one all_gather of the per-candidate fp32 scores per round (and
    nothing inside the T-step loop); every rank then holds the same selection.
  * callable mode — any other `denoise_fn` / `verifier_fn`: the reference's
    serial loop, candidate by candidate, with its exact update rule.

Selection semantics in both modes are the reference's: strict `>` (first maximum
wins, NaN never wins), ZeroOrder moves the pivot only when the round's best beats
the global best (:193-196), PathSearch is the reference's placeholder (perturb
x_T, denoise fully, :307-316) unless `restart=True` is passed (opt-in extension:
true mid-trajectory restart at `injection_step`).

GradientBasedSearch (:343-438) needs autograd through the whole trajectory: it is kept
for arbitrary differentiable callables (callable mode only, the reference's loop); the
kernel sampler is not differentiable and is refused with a clear error.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib
from .._sampler_base import X_T_TAG
from . import verifier as _ver

TAG_X_T = X_T_TAG            # candidate x_T draws
TAG_NEIGHBOR = X_T_TAG + 16  # + iteration
TAG_PATH = X_T_TAG + 8


# ------------------------------------------------------------------ adapter --
class SamplerDenoiser:
    """`denoise_fn` adapter around an its_b200 sampler: the reference's samplers
    do not accept `show_progress`, and its search classes pass the same **kwargs
    to the denoiser and the verifier (search_algorithm.py:71,75)."""

    def __init__(self, sampler, labels: Optional[torch.Tensor] = None, *, max_images: int = 256,
                 seed: Optional[int] = None, step_noise: Optional[torch.Tensor] = None):
        self.sampler = sampler
        self.labels = labels
        self.max_images = int(max_images)
        self.seed = seed
        self.step_noise = step_noise     # [T, B, C, H, W] shared by every candidate (parity runs)
        self._calls = 0                  # callable mode: the i-th call of a search is global candidate i
        self._fallback_seed: Optional[int] = None

    def reset_calls(self) -> None:
        """Callable mode numbers its candidates by call order; a search starts counting at zero."""
        self._calls = 0

    def step_seed(self, shared: Optional[int] = None) -> int:
        """Seed of the per-step Philox noise: the denoiser's own, else the search's shared (broadcast) seed, else
        one draw that then stays fixed for this denoiser (never a fresh one per call: a candidate's trajectory
        must not depend on which call, rank or search round evaluates it)."""
        if self.seed is not None:
            return int(self.seed)
        if shared is not None:
            return int(shared)
        if self._fallback_seed is None:
            self._fallback_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return self._fallback_seed

    def _labels_for(self, n_cand: int, kw_labels) -> Optional[torch.Tensor]:
        lab = kw_labels if kw_labels is not None else self.labels
        if lab is None:
            return None
        return lab.reshape(-1).repeat(n_cand)

    def __call__(self, noise: torch.Tensor, show_progress: bool = False, **kwargs) -> torch.Tensor:
        # the reference draws fresh step noise for every candidate (Diffusion.py:96): candidate i of a serial
        # search uses Philox stream i, exactly the stream population mode gives global candidate i
        i = self._calls
        self._calls += 1
        return self.denoise_candidates(noise.unsqueeze(0), i, labels=kwargs.get("labels"))[0]

    def denoise_candidates(self, cands: torch.Tensor, first_cand: int, *, labels=None,
                           t_start: Optional[int] = None, seed: Optional[int] = None) -> torch.Tensor:
        """cands [n, B, C, H, W] -> images [n, B, C, H, W]; candidate i is global
        candidate first_cand + i (its Philox streams are keyed by that id)."""
        n, B = cands.shape[:2]
        guided = getattr(self.sampler, "guided", False)
        per_call = max(1, self.max_images // B)
        out = torch.empty_like(cands)
        seed = self.step_seed(seed)
        for i0 in range(0, n, per_call):
            i1 = min(n, i0 + per_call)
            x = cands[i0:i1].reshape((i1 - i0) * B, *cands.shape[2:])
            kw: Dict[str, Any] = dict(seed=seed, cand_id0=(first_cand + i0) * B, t_start=t_start)
            if self.step_noise is not None:
                kw["noise"] = self.step_noise.repeat(1, i1 - i0, 1, 1, 1)
            if guided:
                y = self.sampler(x, self._labels_for(i1 - i0, labels), **kw)
            else:
                y = self.sampler(x, **kw)
            out[i0:i1] = y.view(i1 - i0, B, *cands.shape[2:])
        return out


def make_denoise_fn(sampler, labels: Optional[torch.Tensor] = None, **kw) -> SamplerDenoiser:
    """The glue the reference leaves to the user: wrap a sampler as `denoise_fn`."""
    return SamplerDenoiser(sampler, labels, **kw)


def _population_mode(denoise_fn, verifier_fn) -> Optional[Any]:
    if not isinstance(denoise_fn, SamplerDenoiser):
        return None
    owner = getattr(verifier_fn, "__self__", None)
    if isinstance(owner, _ver._KernelVerifier) and getattr(verifier_fn, "__name__", "") == "score":
        if isinstance(owner, _ver.OracleVerifier) and owner.dataset_stats is not None:
            return None
        return owner
    return None


# ---------------------------------------------------------------- utilities --
def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of candidates owned by `rank` (ids stay global, so the
    result does not depend on the number of ranks)."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _gather_scores(local: torch.Tensor, n: int) -> torch.Tensor:
    dist, rank, world = _dist()
    if dist is None:
        return local
    per = -(-n // world)
    buf = torch.full((per,), float("nan"), dtype=torch.float32, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)                       # the path's only collective
    sizes = [_shard(n, r, world)[1] - _shard(n, r, world)[0] for r in range(world)]
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def _shared_seed(seed: Optional[int], device) -> int:
    dist, rank, world = _dist()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    if dist is not None:
        t = torch.tensor([seed], dtype=torch.int64, device=device)
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return seed


def philox_normal(shape: Sequence[int], seed: int, cand_id0: int, tag: int, device, *, base=None,
                  scale: float = 1.0) -> torch.Tensor:
    """[n_units, ...] N(0,1)*scale (+ base broadcast over units) from the library's
    counter-based generator; unit i (a whole candidate tensor) uses stream
    (seed, cand_id0 + i, tag), so a candidate can be regenerated from its id."""
    _lib.require_cuda()
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    n_img = shape[0]
    n_per = out.numel() // n_img
    b = None if base is None else base.to(device=device, dtype=torch.float32).contiguous()
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().its_philox_normal(out.data_ptr(), None if b is None else b.data_ptr(), 1, float(scale),
                                                n_img, n_per, seed, cand_id0, tag, _lib.stream_ptr(out.device)),
                   "its_philox_normal")
    return out


def argmax_first(scores: torch.Tensor) -> Tuple[int, float]:
    """(index, value) of the first maximum; (-1, -inf) when nothing beats -inf."""
    _lib.require_cuda()
    s = scores.detach().to(torch.float32).contiguous()
    idx = torch.empty(1, dtype=torch.int32, device=s.device)
    val = torch.empty(1, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().its_argmax_first(idx.data_ptr(), val.data_ptr(), s.data_ptr(), s.numel(),
                                               _lib.stream_ptr(s.device)), "its_argmax_first")
    return int(idx.item()), float(val.item())


def topk_first(scores: torch.Tensor, k: int) -> Tuple[List[int], List[float]]:
    """The k best (indices, values) under argmax_first's rule (strict '>', first index on ties, NaN never ranks);
    indices are -1 past the number of eligible scores."""
    _lib.require_cuda()
    s = scores.detach().to(torch.float32).contiguous()
    idx = torch.empty(k, dtype=torch.int32, device=s.device)
    val = torch.empty(k, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().its_topk_first(idx.data_ptr(), val.data_ptr(), s.data_ptr(), s.numel(), int(k),
                                             _lib.stream_ptr(s.device)), "its_topk_first")
    return idx.tolist(), val.tolist()


def _score_population(cands_local: torch.Tensor, lo: int, n_total: int, denoise: SamplerDenoiser, ver,
                      kwargs, t_start=None, seed=None) -> torch.Tensor:
    """Denoise + score this rank's block of candidates, return ALL n_total scores.  `seed`: the search's
    shared seed, used for the step noise when the denoiser has none of its own (same on every rank)."""
    B = cands_local.shape[1] if cands_local.numel() else 1
    if cands_local.shape[0] > 0:
        imgs = denoise.denoise_candidates(cands_local, lo, labels=kwargs.get("labels"), t_start=t_start, seed=seed)
        local = ver.score_candidates(imgs.reshape(-1, *imgs.shape[2:]), B)
    else:
        local = torch.empty(0, dtype=torch.float32, device=cands_local.device)
    return _gather_scores(local, n_total)


# ------------------------------------------------------------ RandomSearch --
class RandomSearch:
    """N random x_T candidates; keep the best verifier score (:18-87)."""

    def __init__(self, n_candidates: int = 4):
        self.n_candidates = n_candidates
        self.nfes = 0

    def search(self, noise_shape: Tuple[int, ...], denoise_fn: Callable, verifier_fn: Callable,
               device: str = 'cuda', verbose: bool = True, **kwargs) -> Tuple[torch.Tensor, float]:
        cand_noise = kwargs.pop("candidate_noise", None)   # [n, *noise_shape]: injected candidates
        seed = kwargs.pop("seed", None)
        ver = _population_mode(denoise_fn, verifier_fn)
        n = self.n_candidates
        if isinstance(denoise_fn, SamplerDenoiser):
            denoise_fn.reset_calls()
        if ver is None:
            best_noise, best_score = None, float('-inf')
            for i in range(n):
                noise = cand_noise[i].to(device) if cand_noise is not None else torch.randn(noise_shape, device=device)
                with torch.no_grad():
                    denoised = denoise_fn(noise, show_progress=(i == 0), **kwargs)
                    self.nfes += 1
                score = verifier_fn(denoised, **kwargs)
                if score > best_score:
                    best_score, best_noise = score, noise.clone()
            return best_noise, best_score
        dist, rank, world = _dist()
        dev = torch.device(device)
        seed = _shared_seed(seed, dev)
        self.last_seed = seed
        lo, hi = _shard(n, rank, world)
        B = noise_shape[0]
        if cand_noise is not None:
            local = cand_noise[lo:hi].to(dev, torch.float32)
        else:
            local = philox_normal((hi - lo,) + tuple(noise_shape), seed, lo, TAG_X_T, dev) \
                if hi > lo else torch.empty((0, *noise_shape), device=dev)
        with torch.no_grad():
            scores = _score_population(local, lo, n, denoise_fn, ver, kwargs, seed=seed)
        self.nfes += n
        self.last_scores = scores
        idx, val = argmax_first(scores)
        if idx < 0:
            return None, float('-inf')
        if cand_noise is not None:
            best = cand_noise[idx].to(dev, torch.float32).clone()
        else:  # every rank regenerates the winner from its key: no broadcast needed
            best = philox_normal((1,) + tuple(noise_shape), seed, idx, TAG_X_T, dev)[0]
        self.last_index = idx
        return best, val

    def reset_nfes(self):
        self.nfes = 0


# --------------------------------------------------------- ZeroOrderSearch --
class ZeroOrderSearch:
    """Iterative neighbourhood search around a pivot (:90-235)."""

    def __init__(self, n_neighbors: int = 4, lambda_radius: float = 0.95, n_iterations: int = 10,
                 verbose: bool = False):
        self.n_neighbors = n_neighbors
        self.lambda_radius = lambda_radius
        self.n_iterations = n_iterations
        self.verbose = verbose
        self.nfes = 0

    def _sample_neighbors(self, pivot: torch.Tensor, device: str) -> list:
        """pivot + randn_like(pivot) * (1 - lambda_radius), n_neighbors times (:210-231)."""
        return [pivot + torch.randn_like(pivot) * (1 - self.lambda_radius) for _ in range(self.n_neighbors)]

    def search(self, initial_noise: torch.Tensor, denoise_fn: Callable, verifier_fn: Callable,
               device: str = 'cuda', verbose: Optional[bool] = None,
               **kwargs) -> Tuple[torch.Tensor, float, Dict[str, Any]]:
        perts = kwargs.pop("perturbations", None)    # [iters][K, *shape] injected N(0,1) draws
        seed = kwargs.pop("seed", None)
        ver = _population_mode(denoise_fn, verifier_fn)
        current = initial_noise.clone()
        best_noise, best_score = initial_noise.clone(), float('-inf')
        history: Dict[str, Any] = {'scores': [], 'candidates_per_iter': []}
        K = self.n_neighbors
        radius = 1 - self.lambda_radius
        self.last_index = -1                # global candidate id (round * K + neighbour) of the returned noise
        if isinstance(denoise_fn, SamplerDenoiser):
            denoise_fn.reset_calls()
        if ver is not None:
            dist, rank, world = _dist()
            seed = _shared_seed(seed, initial_noise.device)
            self.last_seed = seed
        for it in range(self.n_iterations):
            if ver is None:
                if perts is not None:
                    neighbors = [current + perts[it][k].to(current) * radius for k in range(K)]
                else:
                    neighbors = self._sample_neighbors(current, device)
                it_scores, it_best, it_best_noise = [], float('-inf'), None
                for k, nb in enumerate(neighbors):
                    with torch.no_grad():
                        denoised = denoise_fn(nb, show_progress=(k == 0), **kwargs)
                        self.nfes += 1
                    score = verifier_fn(denoised, **kwargs)
                    it_scores.append(score)
                    if score > it_best:
                        it_best, it_best_noise, idx = score, nb.clone(), k
            else:
                lo, hi = _shard(K, rank, world)
                B = current.shape[0]
                if perts is not None:
                    local = current.unsqueeze(0) + perts[it][lo:hi].to(current) * radius
                elif hi > lo:
                    local = philox_normal((hi - lo,) + tuple(current.shape), seed, lo, TAG_NEIGHBOR + it,
                                          current.device, base=current.reshape(-1), scale=radius)
                else:
                    local = current.new_empty((0, *current.shape))
                with torch.no_grad():
                    scores = _score_population(local, lo + it * K, K, denoise_fn, ver, kwargs, seed=seed)
                self.nfes += K
                it_scores = [float(s) for s in scores.tolist()]
                idx, it_best = argmax_first(scores)
                it_best_noise = None
                if idx >= 0:
                    if perts is not None:
                        it_best_noise = current + perts[it][idx].to(current) * radius
                    else:
                        it_best_noise = philox_normal((1,) + tuple(current.shape), seed, idx, TAG_NEIGHBOR + it,
                                                      current.device, base=current.reshape(-1),
                                                      scale=radius)[0]
            history['scores'].append(it_scores)
            history['candidates_per_iter'].append(K)
            if it_best > best_score:
                best_score = it_best
                best_noise = it_best_noise.clone()
                current = it_best_noise.clone()
                self.last_index = it * K + idx
        return best_noise, best_score, history

    def reset_nfes(self):
        self.nfes = 0


# -------------------------------------------------------------- PathSearch --
class PathSearch:
    """Search over denoising paths (:238-340).  Default = the reference's
    placeholder: perturb x_T by noise_scale*N(0,1), denoise from T.  With
    `restart=True` (extension) the pivot trajectory is run once down to
    `injection_step`, the perturbation is applied to x_t there and only the
    remaining steps are denoised for every path."""

    def __init__(self, n_paths: int = 4, injection_step: int = 400, noise_scale: float = 0.1,
                 verbose: bool = False):
        self.n_paths = n_paths
        self.injection_step = injection_step
        self.noise_scale = noise_scale
        self.verbose = verbose
        self.nfes = 0

    def search(self, initial_noise: torch.Tensor, denoise_fn: Callable, verifier_fn: Callable,
               timesteps: int = 1000, device: str = 'cuda', verbose: Optional[bool] = None,
               **kwargs) -> Tuple[torch.Tensor, float, Dict[str, Any]]:
        variations = kwargs.pop("variations", None)   # [n_paths, *shape] injected N(0,1) draws
        seed = kwargs.pop("seed", None)
        restart = bool(kwargs.pop("restart", False))
        ver = _population_mode(denoise_fn, verifier_fn)
        best_noise, best_score = initial_noise.clone(), float('-inf')
        history: Dict[str, Any] = {'scores': [], 'injection_points': []}
        P = self.n_paths
        self.last_index = -1                # path index of the returned noise
        if isinstance(denoise_fn, SamplerDenoiser):
            denoise_fn.reset_calls()
        if ver is None:
            if restart:
                raise ValueError("restart=True needs an its_b200 SamplerDenoiser (population mode)")
            for p in range(P):
                var = variations[p].to(initial_noise) if variations is not None else torch.randn_like(initial_noise)
                perturbed = initial_noise + var * self.noise_scale
                with torch.no_grad():
                    denoised = denoise_fn(perturbed, show_progress=(p == 0), **kwargs)
                    self.nfes += 1
                score = verifier_fn(denoised, **kwargs)
                history['scores'].append(score)
                history['injection_points'].append(self.injection_step)
                if score > best_score:
                    best_score, best_noise, self.last_index = score, perturbed.clone(), p
            return best_noise, best_score, history
        dist, rank, world = _dist()
        seed = _shared_seed(seed, initial_noise.device)
        self.last_seed = seed
        lo, hi = _shard(P, rank, world)
        B = initial_noise.shape[0]
        base = initial_noise
        t_start = None
        if restart:
            # pivot trajectory T-1 .. injection_step (exclusive) on every rank, unclipped
            smp = denoise_fn.sampler
            T = smp.T
            if not (0 < self.injection_step < T):
                raise ValueError("injection_step must lie in (0, T)")
            base = _run_prefix(denoise_fn, initial_noise, self.injection_step, kwargs.get("labels"), seed)
            t_start = self.injection_step - 1

        def perturbed(i0, i1):
            if variations is not None:
                return base.unsqueeze(0) + variations[i0:i1].to(base) * self.noise_scale
            if i1 <= i0:
                return base.new_empty((0, *base.shape))
            return philox_normal((i1 - i0,) + tuple(base.shape), seed, i0, TAG_PATH, base.device,
                                 base=base.reshape(-1), scale=self.noise_scale)

        with torch.no_grad():
            scores = _score_population(perturbed(lo, hi), lo, P, denoise_fn, ver, kwargs, t_start=t_start, seed=seed)
        self.nfes += P
        history['scores'] = [float(s) for s in scores.tolist()]
        history['injection_points'] = [self.injection_step] * P
        idx, val = argmax_first(scores)
        if idx >= 0:
            best_score, best_noise, self.last_index = val, perturbed(idx, idx + 1)[0].clone(), idx
        return best_noise, best_score, history

    def reset_nfes(self):
        self.nfes = 0


# ----------------------------------------------------- GradientBasedSearch --
class GradientBasedSearch:
    """Adam ascent on the verifier score through a DIFFERENTIABLE denoise_fn / verifier_fn pair
    (:343-438).  Host-side torch only: the its_b200 kernel sampler has no backward pass, so a
    `SamplerDenoiser` is refused; any autograd-capable callables work as in the reference."""

    def __init__(self, n_iterations: int = 20, lr: float = 0.01, verbose: bool = False):
        self.n_iterations = n_iterations
        self.lr = lr
        self.verbose = verbose
        self.nfes = 0

    def search(self, initial_noise: torch.Tensor, denoise_fn: Callable, verifier_fn: Callable,
               device: str = 'cuda', **kwargs) -> Tuple[torch.Tensor, float, Dict[str, Any]]:
        if isinstance(denoise_fn, SamplerDenoiser):
            raise TypeError("GradientBasedSearch differentiates through denoise_fn; the its_b200 kernel sampler "
                            "is inference-only (use RandomSearch / ZeroOrderSearch / PathSearch with it)")
        noise = initial_noise.clone().requires_grad_(True)
        opt = torch.optim.Adam([noise], lr=self.lr)
        history: Dict[str, Any] = {'scores': [], 'grad_norms': []}
        best_noise, best_score = noise.clone(), float('-inf')
        for it in range(self.n_iterations):
            opt.zero_grad()
            score = verifier_fn(denoise_fn(noise, **kwargs), **kwargs)
            self.nfes += 1
            (-score).backward()                               # maximise the score
            history['grad_norms'].append(noise.grad.norm().item())
            opt.step()                                        # the score recorded below belongs to the pre-step noise,
            value = score.item() if isinstance(score, torch.Tensor) else score
            history['scores'].append(value)
            if value > best_score:                            # ... but, as in the reference, the stored noise is post-step
                best_score, best_noise = value, noise.clone().detach()
            if self.verbose:
                print(f"Gradient-Based Search Iteration {it + 1}/{self.n_iterations}: score {value:.4f}")
        return best_noise, best_score, history

    def reset_nfes(self):
        self.nfes = 0


def _run_prefix(denoise: SamplerDenoiser, x_T: torch.Tensor, stop_step: int, labels, seed=None) -> torch.Tensor:
    """Run the pivot trajectory from T-1 down to `stop_step` inclusive as one device-resident segment
    of the sampler's step graph and return the un-clipped state that step `stop_step - 1` starts from."""
    smp = denoise.sampler
    kw = dict(seed=denoise.step_seed(seed), cand_id0=0, t_stop=stop_step, clip=False)
    if denoise.step_noise is not None:          # parity runs: the pivot's prefix uses the injected Gaussians too
        kw["noise"] = denoise.step_noise
    if getattr(smp, "guided", False):
        lab = (labels if labels is not None else denoise.labels).reshape(-1).to(x_T.device, torch.int64)
        return smp(x_T, lab, **kw)
    return smp(x_T, **kw)
