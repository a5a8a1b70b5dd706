"""Driver-level callers of the sampling path against fixtures produced by
the reference's own functions (tests/golden/make_golden_driver.py)."""
import os

import numpy as np
import pytest
import torch

from tests import cases
from tests.util import build_shell, golden


def test_config_loader_reads_the_reference_yaml_keys(tmp_path):
    from its_b200 import inference as I
    p = tmp_path / "inference_config.yaml"
    p.write_text("hydra:\n  output_subdir: null\nT: 3000\nbeta_1: 1e-4\nbeta_T: 0.02\nimg_size: 256\n"
                 "channel: 128\nchannel_mult: [1, 2, 3, 4]\nattn: [2]\nnum_res_blocks: 2\ndropout: 0.15\n"
                 "batch_size: 64\nmetric_interval: 30\nnrow: 8\n")
    cfg = I.load_config(str(p), ["T=1000", "batch_size=32", "search.algorithm=zero_order", "search.n_neighbors=3"])
    assert "hydra" not in cfg
    assert cfg["T"] == 1000 and cfg["batch_size"] == 32 and cfg["channel_mult"] == [1, 2, 3, 4]
    assert isinstance(cfg["beta_1"], float) and abs(cfg["beta_1"] - 1e-4) < 1e-12
    assert cfg["search"] == {"algorithm": "zero_order", "n_neighbors": 3}
    with pytest.raises(ValueError):
        I.load_config(str(p), ["oops"])


def test_checkpoint_loading_matches_the_reference(tmp_path):
    from its_b200 import inference as I
    g = golden("driver")
    sd = {"module.head.weight": torch.ones(2, 3), "module.time_embedding.timembedding.0.weight": torch.zeros(1000, 8)}
    p = tmp_path / "a.pt"
    torch.save({"state_dict": sd}, str(p))
    got = I.load_checkpoint_state_dict(str(p), torch.device("cpu"))
    assert sorted(got.keys()) == list(g["keys_wrapped"])
    assert (I.detect_checkpoint_T(got) or -1) == int(g["T_table"][0])
    assert (I.detect_checkpoint_T({I.TIME_TABLE_KEY: torch.zeros(512, 128)}) or -1) == int(g["T_linear"][0])
    assert (I.detect_checkpoint_T({"head.weight": torch.zeros(1)}) or -1) == int(g["T_absent"][0])
    torch.save(sd, str(p))                       # bare state dict
    assert sorted(I.load_checkpoint_state_dict(str(p), "cpu").keys()) == list(g["keys_wrapped"])
    with pytest.raises(FileNotFoundError):
        I.load_checkpoint_state_dict(str(tmp_path / "missing.pt"), "cpu")


def test_extended_time_table_is_bit_exact():
    from its_b200 import inference as I
    g = golden("driver")
    assert np.array_equal(I.extended_time_table(600, 900, 16, "interpolate").numpy(), g["table_interp"])
    assert np.array_equal(I.extended_time_table(600, 900, 16, "reinit").numpy(), g["table_reinit"])


def test_metric_points():
    from its_b200 import inference as I
    g = golden("driver")
    assert I.metric_points(20, 6) == [int(t) for t in g["track_hist"][:, 0]]
    assert I.metric_points(10, 5) == [5, 0] and I.metric_points(3, 7) == [0]


def test_table_checkpoint_of_another_T_extends_the_conditional_net(tmp_path):
    """A T=600 table checkpoint sampled at T=900: every other tensor loads, the table is the reference's
    extended one (CPU: construction and loading only)."""
    from its_b200 import inference as I
    from its_b200.DiffusionFreeGuidence import UNet
    cfg = dict(T=900, num_labels=10, channel=64, channel_mult=[1, 2], num_res_blocks=1, dropout=0.0, beta_1=1e-4,
               beta_T=0.02, w=1.8)
    src = UNet(T=600, num_labels=10, ch=64, ch_mult=[1, 2], num_res_blocks=1, dropout=0.0)
    p = tmp_path / "c.pt"
    torch.save({"state_dict": {"module." + k: v for k, v in src.state_dict().items()}}, str(p))
    m = I.create_and_load_model(dict(cfg, checkpoint_path=str(p)), "cpu")
    tab = m.time_embedding.timembedding[0].weight
    assert tuple(tab.shape) == (900, 64)
    assert torch.equal(tab, I.extended_time_table(600, 900, 64))
    assert torch.equal(m.head.weight, src.head.weight)
    assert torch.equal(m.cond_embedding.condEmbedding[1].weight, src.cond_embedding.condEmbedding[1].weight)
    smp = I.create_sampler(m, cfg, "cpu")
    assert smp.w == 1.8 and smp.T == 900
    assert m.residual_fp16 is True                       # auto: a checkpoint was loaded
    assert I.create_and_load_model(dict(cfg, checkpoint_path=None), "cpu").residual_fp16 is False
    assert I.create_and_load_model(dict(cfg, checkpoint_path=str(p), residual_fp16=False), "cpu").residual_fp16 is False


class _FakeFID:
    def extract_features_from_tensor(self, x01):
        return x01.flatten(1)[:, :6].double()

    def calculate_frechet_distance(self, mu_r, s_r, mu_f, s_f):
        return float((mu_r - mu_f).pow(2).sum() + (s_r - s_f).diagonal().abs().sum())


class _FakeIS:
    def compute_is(self, x01):
        return float(x01.mean()), float(x01.std())


@pytest.mark.gpu
def test_metrics_tracking_matches_the_reference_loop(cuda_dev, built_lib):
    """sample_with_metrics_tracking against Diffusion/Train.py:25-166 run on the same weights, x_T and
    injected noise with stand-in calculators: same metric points, metrics within the 16-bit tolerance,
    final samples within 2e-2."""
    from its_b200 import inference as I
    from its_b200.Diffusion import GaussianDiffusionSampler
    g = golden("driver")
    cfg = cases.SAMPLER_CASES["u_small_T20"]
    net, _ = build_shell(cfg, cuda_dev)
    smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"]).to(cuda_dev)
    smp.print_steps = False
    x_T, noise, _ = cases.sampler_inputs(cfg)
    real = torch.from_numpy(g["track_real"]).to(cuda_dev)
    x0, hist = I.sample_with_metrics_tracking(smp, x_T.to(cuda_dev), _FakeFID(), _FakeIS(), None, real,
                                              torch.zeros(2, 4), metric_interval=6, device=str(cuda_dev),
                                              noise=noise.to(cuda_dev))
    ref_hist = g["track_hist"]
    assert [h[0] for h in hist] == [int(t) for t in ref_hist[:, 0]]
    for (t, fid, isv, clip), r in zip(hist, ref_hist):
        assert abs(fid - r[1]) <= 2e-2 * max(1.0, abs(r[1])), (t, fid, r[1])
        assert abs(isv - r[2]) <= 5e-3, (t, isv, r[2])
        assert np.isnan(clip) and np.isnan(r[3])
    err = (x0.cpu() - torch.from_numpy(g["track_x0"])).abs().max().item()
    assert err <= 2e-2, err
    # a trajectory cut into segments is the uncut trajectory, bit for bit
    whole = smp(x_T.to(cuda_dev), noise=noise.to(cuda_dev))
    assert torch.equal(whole, x0)
    x0p, _ = I.sample_with_metrics_tracking(smp, x_T.to(cuda_dev), metric_interval=1, seed=7)
    assert torch.equal(x0p, smp(x_T.to(cuda_dev), seed=7))


@pytest.mark.gpu
def test_config_driven_search_and_grid(cuda_dev, built_lib, tmp_path):
    from its_b200 import inference as I
    cfg = dict(T=8, beta_1=1e-4, beta_T=0.02, img_size=16, channel=64, channel_mult=[1, 2], attn=[1], num_res_blocks=1,
               dropout=0.0, batch_size=2, checkpoint_path=None, nrow=2,
               search=dict(algorithm="random", n_candidates=5, verifier="oracle"))
    model = I.create_and_load_model(cfg, cuda_dev)
    smp = I.create_sampler(model, cfg, cuda_dev)
    smp.print_steps = False
    best, score, images = I.run_search(smp, cfg, cuda_dev, seed=3)
    assert tuple(best.shape) == (2, 3, 16, 16) and tuple(images.shape) == (2, 3, 16, 16)
    best2, score2, _ = I.run_search(smp, cfg, cuda_dev, seed=3)
    assert torch.equal(best, best2) and score == score2          # keyed streams: reproducible
    for algo in ("zero_order", "path"):
        c2 = dict(cfg, search=dict(algorithm=algo, n_neighbors=2, n_iterations=2, n_paths=2, injection_step=4,
                                   verifier="aesthetic"))
        b, s, im = I.run_search(smp, c2, cuda_dev, seed=3)
        assert b is not None and np.isfinite(s) and im.abs().max().item() <= 1.0
    out = I.save_image_grid(images, str(tmp_path / "grid" / "s.png"), nrow=2)
    assert os.path.getsize(out) > 100


def test_repo_config_file_parses():
    from its_b200 import inference as I
    from tests.conftest import ROOT
    cfg = I.load_config(os.path.join(ROOT, "config", "inference_config.yaml"), ["T=8", "search.n_candidates=3"])
    assert cfg["T"] == 8 and cfg["channel_mult"] == [1, 2, 3, 4] and cfg["checkpoint_path"] is None
    assert cfg["search"]["n_candidates"] == 3 and cfg["search"]["algorithm"] == "random"


@pytest.mark.gpu
def test_cli_main_end_to_end(cuda_dev, built_lib, tmp_path, capsys):
    """`python -m its_b200.inference <yaml> key=value ...` from the repo's config file: search, then sampling with
    metrics tracking, image grids written."""
    from its_b200 import inference as I
    from tests.conftest import ROOT
    cfg_path = os.path.join(ROOT, "config", "inference_config.yaml")
    small = ["T=8", "img_size=16", "channel=64", "channel_mult=[1,2]", "num_res_blocks=1", "dropout=0.0", "batch_size=2",
             f"sampled_images_save_dir={tmp_path}", "nrow=2"]
    assert I.main([cfg_path] + small + ["search.n_candidates=4"]) == 0
    assert "search: best score" in capsys.readouterr().out
    assert I.main([cfg_path] + small + ["search=null", "metric_interval=3"]) == 0
    assert "metric points" in capsys.readouterr().out
    assert any(f.endswith(".png") for f in os.listdir(tmp_path))
