"""Generate tests/golden/driver.npz by running the REFERENCE's driver-level functions
(SURVEY.md §8f rows 1 and 3): `sample_with_metrics_tracking` (Diffusion/Train.py:25-166) and
`reinitialize_time_embedding` / `detect_checkpoint_T` / `load_checkpoint_state_dict`
(abstract_metrics_from_pretrained_ddpm.py:126-262).

Build container only (needs /root/reference; CPU):   python tests/golden/make_golden_driver.py

The reference files are executed unmodified; modules that are absent here and irrelevant to these
functions (matplotlib, hydra, omegaconf) are pre-seeded as empty stubs, as SURVEY.md §8c describes.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ITS_REF_DIR", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import ddpm_oracle as O  # noqa: E402
from tests import cases  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot", "omegaconf", "hydra"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["omegaconf"].DictConfig = dict
sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {"to_container": staticmethod(lambda c, **k: dict(c))})
sys.modules["hydra"].main = lambda **kw: (lambda fn: fn)
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    import Diffusion as ref_pkg                      # noqa: E402  (the reference package)
    from Diffusion.Train import sample_with_metrics_tracking as ref_track  # noqa: E402
    spec = importlib.util.spec_from_file_location("ref_driver", os.path.join(REF, "abstract_metrics_from_pretrained_ddpm.py"))
    ref_driver = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_driver)


class FakeFID:
    """Deterministic stand-in with the reference calculators' protocol (utils/metrics.py)."""

    def extract_features_from_tensor(self, x01):
        return x01.flatten(1)[:, :6].double()

    def calculate_frechet_distance(self, mu_r, s_r, mu_f, s_f):
        return float((mu_r - mu_f).pow(2).sum() + (s_r - s_f).diagonal().abs().sum())


class FakeIS:
    def compute_is(self, x01):
        return float(x01.mean()), float(x01.std())


def tracking_case():
    cfg = cases.SAMPLER_CASES["u_small_T20"]
    m = ref_pkg.Model.UNet(T=cfg["T"], ch=cfg["ch"], ch_mult=cfg["ch_mult"], attn=cfg["attn"],
                           num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"])
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(O.synth_state_dict(shapes, cfg["weight_seed"]), strict=True)
    m.eval()
    smp = ref_pkg.GaussianDiffusionSampler(m, cfg["beta_1"], cfg["beta_T"], cfg["T"])
    x_T, noise, _ = cases.sampler_inputs(cfg)
    real = torch.from_numpy(np.random.default_rng(9).random((5, 6))).double()
    calls = {"t": cfg["T"] - 1}
    orig = torch.randn_like

    def injected(x, *a, **k):                # draw of time_step t (T-1 ... 1), in loop order
        z = noise[calls["t"]]
        calls["t"] -= 1
        return z.to(x)

    torch.randn_like = injected
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            x0, hist = ref_track(smp, x_T, FakeFID(), FakeIS(), None, real, torch.zeros(2, 4), metric_interval=6, device="cpu")
    finally:
        torch.randn_like = orig
    return x0.numpy(), np.array(hist, dtype=np.float64), real.numpy()


def table_cases():
    out = {}
    for tag, strategy in (("interp", "interpolate"), ("reinit", "reinit")):
        holder = torch.nn.Module()
        holder.time_embedding = torch.nn.Module()
        holder.time_embedding.timembedding = torch.nn.Sequential(
            torch.nn.Embedding(600, 16), torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 32))
        with contextlib.redirect_stdout(io.StringIO()):
            ref_driver.reinitialize_time_embedding(holder, 600, 900, {"time_embedding_strategy": strategy}, torch.device("cpu"))
        out[tag] = holder.time_embedding.timembedding[0].weight.detach().numpy().copy()
    return out


def checkpoint_cases():
    res = {}
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
        sd = {"module.head.weight": torch.ones(2, 3), "module.time_embedding.timembedding.0.weight": torch.zeros(1000, 8)}
        p1 = os.path.join(d, "a.pt")
        torch.save({"state_dict": sd}, p1)
        got = ref_driver.load_checkpoint_state_dict(p1, torch.device("cpu"))
        res["keys_wrapped"] = np.array(sorted(got.keys()))
        res["T_table"] = np.array([ref_driver.detect_checkpoint_T(got) or -1])
        res["T_linear"] = np.array([ref_driver.detect_checkpoint_T({"time_embedding.timembedding.0.weight": torch.zeros(512, 128)}) or -1])
        res["T_absent"] = np.array([ref_driver.detect_checkpoint_T({"head.weight": torch.zeros(1)}) or -1])
    return res


if __name__ == "__main__":
    x0, hist, real = tracking_case()
    tabs = table_cases()
    ck = checkpoint_cases()
    np.savez_compressed(os.path.join(HERE, "driver.npz"), track_x0=x0, track_hist=hist, track_real=real,
                        table_interp=tabs["interp"], table_reinit=tabs["reinit"], **ck)
    print("driver.npz:", {k: v.shape for k, v in dict(track_x0=x0, track_hist=hist, **tabs).items()}, ck)
