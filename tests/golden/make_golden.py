"""Generate tests/golden/*.npz by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, CPU is enough):

    python tests/golden/make_golden.py

The reference is imported unmodified (Diffusion/*.py by file path because
`import Diffusion` pulls matplotlib through Train.py).  Weights
come from oracle.ddpm_oracle.synth_state_dict (numpy RNG, so the GPU box can
rebuild them without torch-RNG coupling) loaded into the reference modules with
load_state_dict(strict=True).  Only inputs' seeds and the reference OUTPUTS are
stored, which keeps the fixtures small.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ITS_REF_DIR", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import ddpm_oracle as O  # noqa: E402
from tests import cases  # noqa: E402


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_diffusion = load_by_path("ref_diffusion", "Diffusion/Diffusion.py")
ref_model = load_by_path("ref_model", "Diffusion/Model.py")
ref_cdiffusion = load_by_path("ref_cdiffusion", "DiffusionFreeGuidence/DiffusionCondition.py")
ref_cmodel = load_by_path("ref_cmodel", "DiffusionFreeGuidence/ModelCondition.py")
with contextlib.redirect_stdout(io.StringIO()):
    ref_search = load_by_path("ref_search", "search/search_algorithm.py")
    ref_verifier = load_by_path("ref_verifier", "search/verifier.py")


def build_ref_model(cfg):
    if cfg["kind"] == "uncond":
        m = ref_model.UNet(T=cfg["T"], ch=cfg["ch"], ch_mult=cfg["ch_mult"], attn=cfg["attn"],
                           num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"])
    else:
        m = ref_cmodel.UNet(T=cfg["T"], num_labels=cfg["num_labels"], ch=cfg["ch"], ch_mult=cfg["ch_mult"],
                            num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"])
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = O.synth_state_dict(shapes, cfg["weight_seed"])
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m, sd, shapes


def ref_sample(sampler, x_T, noise, labels=None):
    """Drive the reference sampler with injected noise by patching randn_like
    (the reference draws one tensor per step for time_step = T-1 .. 1)."""
    T = sampler.T
    calls = {"n": 0}
    orig = torch.randn_like

    def fake(x, *a, **k):
        step = T - 1 - calls["n"]
        calls["n"] += 1
        return noise[step]

    torch.randn_like = fake
    try:
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            out = sampler(x_T) if labels is None else sampler(x_T, labels)
    finally:
        torch.randn_like = orig
    assert calls["n"] == T - 1
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrays.items()})


def main():
    torch.set_num_threads(8)
    # ---- schedule known answers (Diffusion.py:57-65,76) ----
    kat = {}
    for T, bT in ((1000, 0.02), (2000, 0.02), (3000, 0.028)):
        s = ref_diffusion.GaussianDiffusionSampler(torch.nn.Identity(), 1e-4, bT, T)
        var = torch.cat([s.posterior_var[1:2], s.betas[1:]])
        kat[f"T{T}_betas"] = s.betas.numpy()
        kat[f"T{T}_coeff1"] = s.coeff1.numpy()
        kat[f"T{T}_coeff2"] = s.coeff2.numpy()
        kat[f"T{T}_posterior_var"] = s.posterior_var.numpy()
        kat[f"T{T}_var"] = var.numpy()
    save("schedule", **kat)

    # ---- UNet forward cases ----
    for name, cfg in cases.FORWARD_CASES.items():
        m, sd, shapes = build_ref_model(cfg)
        x, t, labels = cases.forward_inputs(cfg)
        with torch.no_grad():
            y = m(x, t) if labels is None else m(x, t, labels)
            y_or = O.unet_forward(sd, x, t, labels)
        err = (y - y_or).abs().max().item()
        print(f"{name}: |ref| max {y.abs().max():.4f}  oracle-vs-ref max abs {err:.3e}")
        assert err < 1e-4 * max(1.0, y.abs().max().item()), "oracle restatement disagrees with the reference"
        save("fwd_" + name, eps=y.numpy())

    # ---- sampler cases (injected noise) ----
    for name, cfg in cases.SAMPLER_CASES.items():
        m, sd, shapes = build_ref_model(cfg)
        x_T, noise, labels = cases.sampler_inputs(cfg)
        if cfg["kind"] == "uncond":
            smp = ref_diffusion.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"])
        else:
            smp = ref_cdiffusion.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"], w=cfg["w"])
        x0 = ref_sample(smp, x_T, noise, labels)
        sched = O.schedule(cfg["beta_1"], cfg["beta_T"], cfg["T"])
        with torch.no_grad():
            x0_or = O.sample(sd, sched, x_T, lambda s: noise[s], labels, cfg.get("w", 0.0))
        err = (x0 - x0_or).abs().max().item()
        print(f"{name}: sampler oracle-vs-ref max abs {err:.3e}; saturated {(x0.abs() == 1).float().mean():.3f}")
        assert err < 2e-4
        scores = {k: f(x0) for k, f in O.VERIFIERS.items()}
        ref_scores = {
            "oracle": ref_verifier.OracleVerifier().score(x0),
            "aesthetic": ref_verifier.AestheticPredictor(device="cpu").score(x0),
            "self_supervised": ref_verifier.SelfSupervisedVerifier().score(x0),
        }
        for k in scores:
            assert abs(scores[k] - ref_scores[k]) < 1e-6, (k, scores[k], ref_scores[k])
        save("smp_" + name, x0=x0.numpy(), **{"score_" + k: v for k, v in ref_scores.items()})

    # ---- verifier known answers on fixed images ----
    imgs = cases.verifier_images()
    vs = {}
    for i, im in enumerate(imgs):
        vs[f"oracle_{i}"] = ref_verifier.OracleVerifier().score(im)
        vs[f"aesthetic_{i}"] = ref_verifier.AestheticPredictor(device="cpu").score(im)
        vs[f"self_supervised_{i}"] = ref_verifier.SelfSupervisedVerifier().score(im)
        assert abs(vs[f"oracle_{i}"] - O.oracle_verifier_score(im)) < 1e-6
        assert abs(vs[f"aesthetic_{i}"] - O.aesthetic_score(im)) < 1e-6
        a, b = vs[f"self_supervised_{i}"], O.self_supervised_score(im)
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-6
    save("verifier", **vs)

    # ---- search cases: the reference's own search classes over its own sampler ----
    for name, cfg in cases.SEARCH_CASES.items():
        m, sd, shapes = build_ref_model(cfg)
        sched_noise = cases.search_noise(cfg)
        if cfg["kind"] == "uncond":
            smp = ref_diffusion.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"])
        else:
            smp = ref_cdiffusion.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"], w=cfg["w"])
        labels = cases.search_labels(cfg)

        def denoise_fn(noise, show_progress=False, **kw):
            return ref_sample(smp, noise, sched_noise, labels)

        ver = {"oracle": ref_verifier.OracleVerifier().score,
               "aesthetic": ref_verifier.AestheticPredictor(device="cpu").score,
               "self_supervised": ref_verifier.SelfSupervisedVerifier().score}[cfg["verifier"]]

        def verifier_fn(images, **kw):
            return ver(images)

        shape = tuple(cfg["noise_shape"])
        out = {}
        # random search: candidates drawn by the reference from torch's global RNG
        torch.manual_seed(cfg["search_seed"])
        rs = ref_search.RandomSearch(n_candidates=cfg["n_candidates"])
        best_noise, best_score = rs.search(shape, denoise_fn, verifier_fn, device="cpu", verbose=False)
        torch.manual_seed(cfg["search_seed"])
        cands = [torch.randn(shape) for _ in range(cfg["n_candidates"])]
        scores = [verifier_fn(denoise_fn(c)) for c in cands]
        idx = int(np.argmax([(-np.inf if np.isnan(s) else s) for s in scores]))
        assert torch.equal(best_noise, cands[idx]) and scores[idx] == best_score and rs.nfes == cfg["n_candidates"]
        out.update(rs_candidates=torch.stack(cands).numpy(), rs_scores=np.array(scores), rs_best_index=idx,
                   rs_best_score=best_score)
        # zero-order search: pivot + perturbations drawn by the reference
        torch.manual_seed(cfg["search_seed"] + 1)
        init = torch.randn(shape)
        zo = ref_search.ZeroOrderSearch(n_neighbors=cfg["zo_neighbors"], lambda_radius=0.95,
                                        n_iterations=cfg["zo_iterations"])
        torch.manual_seed(cfg["search_seed"] + 2)
        zo_noise, zo_score, zo_hist = zo.search(init, denoise_fn, verifier_fn, device="cpu", verbose=False)
        torch.manual_seed(cfg["search_seed"] + 2)
        # replay the draws: each iteration draws n_neighbors randn_like(pivot) tensors up front;
        # the sampler's own noise is injected so nothing else touches the RNG
        perts = [[torch.randn(shape) for _ in range(cfg["zo_neighbors"])] for _ in range(cfg["zo_iterations"])]
        o_noise, o_score, o_hist = O.zero_order_search(init, perts, 0.95, denoise_fn, verifier_fn)
        assert torch.allclose(o_noise, zo_noise) and o_score == zo_score and o_hist["scores"] == zo_hist["scores"]
        out.update(zo_init=init.numpy(), zo_perts=torch.stack([torch.stack(p) for p in perts]).numpy(),
                   zo_scores=np.array(zo_hist["scores"]), zo_best_score=zo_score, zo_best_noise=zo_noise.numpy())
        # path search
        ps = ref_search.PathSearch(n_paths=cfg["n_paths"], injection_step=cfg["T"] // 2, noise_scale=0.1)
        torch.manual_seed(cfg["search_seed"] + 3)
        ps_noise, ps_score, ps_hist = ps.search(init, denoise_fn, verifier_fn, timesteps=cfg["T"], device="cpu",
                                                verbose=False)
        torch.manual_seed(cfg["search_seed"] + 3)
        variations = [torch.randn(shape) for _ in range(cfg["n_paths"])]
        o_noise, o_score, o_hist = O.path_search(init, variations, 0.1, cfg["T"] // 2, denoise_fn, verifier_fn)
        assert torch.allclose(o_noise, ps_noise) and o_score == ps_score and o_hist["scores"] == ps_hist["scores"]
        out.update(ps_variations=torch.stack(variations).numpy(), ps_scores=np.array(ps_hist["scores"]),
                   ps_best_score=ps_score, ps_best_noise=ps_noise.numpy())
        save("search_" + name, **out)


if __name__ == "__main__":
    main()
