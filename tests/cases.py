"""Shared definitions of the parity cases (used by tests/golden/make_golden.py in
the build container and by the tests everywhere).  Inputs come from numpy RNGs so
both sides can rebuild them from a seed."""
from __future__ import annotations

import numpy as np
import torch

U_SMALL = dict(kind="uncond", T=50, ch=64, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0,
               weight_seed=11, img=32, B=2, input_seed=101)
U_3LVL = dict(kind="uncond", T=50, ch=64, ch_mult=[1, 2, 2], attn=[1], num_res_blocks=1, dropout=0.0,
              weight_seed=12, img=32, B=8, input_seed=102)
U_A = dict(kind="uncond", T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1], num_res_blocks=2, dropout=0.15,
           weight_seed=13, img=32, B=2, input_seed=103)
U_E = dict(kind="uncond", T=2000, ch=128, ch_mult=[1, 2, 3, 4], attn=[2], num_res_blocks=2, dropout=0.15,
           weight_seed=14, img=64, B=2, input_seed=104)
C_SMALL = dict(kind="cond", T=50, num_labels=10, ch=64, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0,
               weight_seed=21, img=16, B=2, input_seed=201)
C_C = dict(kind="cond", T=1000, num_labels=10, ch=128, ch_mult=[1, 2, 3, 4], num_res_blocks=2, dropout=0.15,
           weight_seed=22, img=32, B=2, input_seed=202)

FORWARD_CASES = {"u_small": U_SMALL, "u_3lvl": U_3LVL, "u_A": U_A, "u_E": U_E, "c_small": C_SMALL, "c_C": C_C}

SAMPLER_CASES = {
    "u_small_T20": dict(U_SMALL, T=20, beta_1=1e-4, beta_T=0.02, noise_seed=301),
    "c_small_T20": dict(C_SMALL, T=20, beta_1=1e-4, beta_T=0.02, w=1.8, noise_seed=302),
}

SEARCH_CASES = {
    "u_search": dict(kind="uncond", T=8, ch=64, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0,
                     weight_seed=31, beta_1=1e-4, beta_T=0.02, noise_shape=[2, 3, 16, 16], noise_seed=401,
                     search_seed=501, n_candidates=6, zo_neighbors=3, zo_iterations=2, n_paths=3,
                     verifier="oracle"),
    "c_search": dict(kind="cond", T=6, num_labels=10, ch=64, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0,
                     weight_seed=32, beta_1=1e-4, beta_T=0.02, w=1.8, noise_shape=[2, 3, 16, 16],
                     noise_seed=402, search_seed=502, n_candidates=5, zo_neighbors=2, zo_iterations=2,
                     n_paths=2, verifier="aesthetic"),
}


# Full-length / full-width fixtures (tests/golden/make_golden_long.py, make_golden_long2.py)
LONG_CASES = {
    "u_A_T1000": dict(U_A, T=1000, beta_1=1e-4, beta_T=0.02, B=2, input_seed=601, noise_seed=602, weight_seed=61,
                      keep_at=(900, 500, 100)),
    "c_C_T1000": dict(C_C, T=1000, beta_1=1e-4, beta_T=0.02, w=1.8, B=2, input_seed=611, noise_seed=612,
                      weight_seed=62, keep_at=(900, 500, 100)),
    "u_E_T2000": dict(U_E, T=2000, beta_1=1e-4, beta_T=0.02, B=1, input_seed=621, noise_seed=622, weight_seed=63,
                      keep_at=(1800, 1000, 200)),
    # the reference's own initialisers under torch.manual_seed(init_seed) (no synthetic weights)
    "u_A_refinit": dict(U_A, T=1000, beta_1=1e-4, beta_T=0.02, B=2, input_seed=631, noise_seed=632, init_seed=0,
                        keep_at=(900, 500, 100)),
    # random search at config A's real width: 64 single-image candidates, T = 50
    "u_A_search64": dict(U_A, T=50, beta_1=1e-4, beta_T=0.02, weight_seed=64, noise_shape=[1, 3, 32, 32],
                         noise_seed=641, cand_seed=642, n_candidates=64, verifier="oracle"),
}


def _randn(seed, shape):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape).astype(np.float32))


def forward_inputs(cfg):
    rng = np.random.default_rng(cfg["input_seed"])
    B, S = cfg["B"], cfg["img"]
    x = torch.from_numpy(rng.standard_normal((B, 3, S, S)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, cfg["T"], size=(B,)).astype(np.int64))
    labels = None
    if cfg["kind"] == "cond":
        labels = torch.from_numpy(rng.integers(0, cfg["num_labels"] + 1, size=(B,)).astype(np.int64))
        labels[0] = 0  # exercise the null class
    return x, t, labels


def sampler_inputs(cfg):
    B, S, T = cfg["B"], cfg["img"], cfg["T"]
    x_T = _randn(cfg["input_seed"], (B, 3, S, S))
    noise = _randn(cfg["noise_seed"], (T, B, 3, S, S))
    labels = None
    if cfg["kind"] == "cond":
        labels = torch.tensor([1 + (i % cfg["num_labels"]) for i in range(B)], dtype=torch.long)
    return x_T, noise, labels


def search_noise(cfg):
    return _randn(cfg["noise_seed"], (cfg["T"],) + tuple(cfg["noise_shape"]))


def search_candidates(cfg):
    """Candidate noises [n_candidates, *noise_shape] from a numpy seed (injected into the searches)."""
    return _randn(cfg["cand_seed"], (cfg["n_candidates"],) + tuple(cfg["noise_shape"]))


def search_labels(cfg):
    if cfg["kind"] != "cond":
        return None
    B = cfg["noise_shape"][0]
    return torch.tensor([1 + (i % cfg["num_labels"]) for i in range(B)], dtype=torch.long)


def verifier_images():
    """Fixed image batches in [-1, 1] and [0, 1] (the aesthetic score branches on
    min < 0), one single-image batch (self-supervised score is NaN there)."""
    rng = np.random.default_rng(777)
    a = np.clip(rng.standard_normal((4, 3, 32, 32)) * 0.6, -1, 1).astype(np.float32)
    b = rng.random((3, 3, 32, 32)).astype(np.float32)
    c = np.clip(rng.standard_normal((1, 3, 64, 64)), -1, 1).astype(np.float32)
    d = np.clip(rng.standard_normal((8, 3, 64, 64)) * 2.0, -1, 1).astype(np.float32)
    return [torch.from_numpy(v) for v in (a, b, c, d)]
