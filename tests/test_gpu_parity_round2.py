"""GPU parity, round 2: the gaps VERDICT round 1 listed.

* the guided sampler's public seam p_mean_variance(x_t, t, labels) against the reference's own outputs;
* full-length trajectories of config C (classifier-free guidance, w = 1.8, T = 1000) and config E (64x64, T = 2000)
  against fixtures produced by running the unmodified reference (tests/golden/make_golden_long2.py);
* BASELINE north_star's literal case: the reference's OWN random initialisation (torch.manual_seed + constructor);
* random search over 64 candidates at config A's real width against the reference's scores and selection;
* search over paths with a mid-trajectory restart against the oracle's restatement.

Tolerances (north_star): samples within max-abs 2e-2 (16-bit operands), verifier scores within 1e-3, selected index
exact whenever the top-2 margin exceeds the score tolerance."""
import math

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from tests import cases
from tests.util import build_shell, golden, rel_err, rms_err

pytestmark = pytest.mark.gpu

SAMPLE_TOL = 2e-2
SCORE_TOL = 1e-3


def _sampler(cfg, net, dev):
    if cfg["kind"] == "uncond":
        from its_b200.Diffusion import GaussianDiffusionSampler
        smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"])
    else:
        from its_b200.DiffusionFreeGuidence import GaussianDiffusionSampler
        smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"], w=cfg["w"])
    smp = smp.to(dev)
    smp.print_steps = False
    return smp


def test_guided_p_mean_variance_seam_vs_reference(cuda_dev):
    """DiffusionCondition.py:79-87 through the shell's own method, exactly as an external loop would call it:
    two UNet evaluations (labels, null labels), (1+w) eps - w nonEps, posterior mean; variance table entry."""
    cfg = cases.SAMPLER_CASES["c_small_T20"]
    g = golden("seam_c_small")
    net, _ = build_shell(cfg, cuda_dev)
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, labels = cases.sampler_inputs(cfg)
    for ts in (cfg["T"] - 1, cfg["T"] // 2, 0):
        x_t = (x_T + 0.25 * noise[ts]).to(cuda_dev)
        t = torch.full((x_T.shape[0],), ts, dtype=torch.long, device=cuda_dev)
        mean, var = smp.p_mean_variance(x_t=x_t, t=t, labels=labels.to(cuda_dev))
        assert var.shape == (x_T.shape[0], 1, 1, 1)
        assert torch.equal(var.cpu(), torch.from_numpy(g[f"var_{ts}"]))          # fp64 table -> fp32, exact
        err = (mean.cpu() - torch.from_numpy(g[f"mean_{ts}"])).abs().max().item()
        print("guided seam, step %d: max abs err of the posterior mean %.2e" % (ts, err))
        assert err <= SAMPLE_TOL
        # the fused guidance mix of the sampler's own step computes the same mean
        fused = smp(x_t, labels.to(cuda_dev), noise=torch.zeros((cfg["T"],) + tuple(x_t.shape), device=cuda_dev),
                    t_start=ts, t_stop=ts, clip=False)
        assert (fused - mean).abs().max().item() <= 1e-5 * max(1.0, mean.abs().max().item())


@pytest.mark.parametrize("name,fixture,rel_tol,rms_tol", [("c_C_T1000", "smp_c_C_T1000", 2e-2, 1e-2),
                                                          ("u_E_T2000", "smp_u_E_T2000", 1e-2, 8e-3)])
def test_full_length_trajectory_configs_C_and_E_vs_reference(cuda_dev, name, fixture, rel_tol, rms_tol):
    """All steps of config C (guided, w = 1.8, T = 1000, B = 2) and config E (64x64, T = 2000, B = 1) at their real
    widths against the reference sampler on the same synthetic O(1) weights, x_T, labels and injected noise.
    The un-clipped state is compared at four checkpoints (relative: an untrained net drives |x| to ~1e3 and
    beyond before the final clip, SURVEY.md section 8c probe 7), the clipped samples on every pixel outside the
    arithmetic noise of the clip boundary."""
    cfg = cases.LONG_CASES[name]
    g = golden(fixture)
    net, _ = build_shell(cfg, cuda_dev)
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, labels = cases.sampler_inputs(cfg)
    x, noise = x_T.to(cuda_dev), noise.to(cuda_dev)
    lab = None if labels is None else labels.to(cuda_dev)
    first = cfg["T"] - 1
    for stop in tuple(cfg["keep_at"]) + (0,):
        kw = dict(noise=noise, t_start=first, t_stop=stop, clip=False)
        x = smp(x, lab, **kw) if lab is not None else smp(x, **kw)
        ref = torch.from_numpy(g["x0_preclip"] if stop == 0 else g[f"x_after_{stop}"]).to(cuda_dev)
        print("%s after step %d: |x|max %.3e rel err %.2e rms err %.2e" % (name, stop, ref.abs().max().item(),
                                                                          rel_err(x, ref), rms_err(x, ref)))
        assert rel_err(x, ref) < rel_tol, (stop, rel_err(x, ref))
        assert rms_err(x, ref) < rms_tol, (stop, rms_err(x, ref))
        first = stop - 1
    ref_pre = torch.from_numpy(g["x0_preclip"]).to(cuda_dev)
    d = (torch.clip(x, -1, 1) - torch.from_numpy(g["x0"]).to(cuda_dev)).abs()
    well = ref_pre.abs() > 4 * rel_tol * ref_pre.abs().max()      # outside the error band around the clip boundary
    assert d[well].max().item() <= SAMPLE_TOL
    whole = smp(x_T.to(cuda_dev), lab, noise=noise) if lab is not None else smp(x_T.to(cuda_dev), noise=noise)
    assert torch.equal(whole, torch.clip(x, -1, 1))             # the one-call forward is the same trajectory


def test_reference_own_random_init_trajectory(cuda_dev):
    """north_star's literal case: torch.manual_seed(0) + the constructor (the shell draws the reference's parameters
    bit for bit, tests/test_oracle_vs_golden.py::test_shell_constructor_reproduces_the_reference_init), all 1000
    steps of config A, injected noise.  With the zero-gain output initialisers eps ~ 3e-5 and the fused step is
    the reference's own expression order, so EVERY pixel of the clipped samples agrees within 2e-2 — no carve-out."""
    cfg = cases.LONG_CASES["u_A_refinit"]
    g = golden("smp_u_A_refinit")
    from its_b200.Diffusion import UNet
    torch.manual_seed(cfg["init_seed"])
    net = UNet(T=cfg["T"], ch=cfg["ch"], ch_mult=cfg["ch_mult"], attn=cfg["attn"],
               num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"]).eval().to(cuda_dev)
    x, t, _ = cases.forward_inputs(cfg)
    eps = net(x.to(cuda_dev), t.to(cuda_dev)).cpu()
    ref_eps = torch.from_numpy(g["eps"])
    print("reference-init eps: max |ref| %.3e, max abs err %.2e" % (ref_eps.abs().max().item(),
                                                                   (eps - ref_eps).abs().max().item()))
    # the zero-gain output initialisers put eps at ~3e-5: what matters for the trajectory is the absolute error
    assert (eps - ref_eps).abs().max().item() <= 5e-6
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, _ = cases.sampler_inputs(cfg)
    x, noise = x_T.to(cuda_dev), noise.to(cuda_dev)
    first = cfg["T"] - 1
    for stop in tuple(cfg["keep_at"]) + (0,):
        x = smp(x, noise=noise, t_start=first, t_stop=stop, clip=False)
        ref = torch.from_numpy(g["x0_preclip"] if stop == 0 else g[f"x_after_{stop}"]).to(cuda_dev)
        print("ref-init after step %d: |x|max %.3e rel err %.2e" % (stop, ref.abs().max().item(), rel_err(x, ref)))
        assert rel_err(x, ref) < 1e-5, (stop, rel_err(x, ref))
        first = stop - 1
    d = (torch.clip(x, -1, 1) - torch.from_numpy(g["x0"]).to(cuda_dev)).abs()
    print("ref-init clipped samples: max abs err %.2e over all %d pixels" % (d.max().item(), d.numel()))
    assert d.max().item() <= SAMPLE_TOL                          # every pixel, nothing excused
    assert torch.equal(smp(x_T.to(cuda_dev), noise=noise), torch.clip(x, -1, 1))


def test_random_search_64_candidates_config_A_vs_reference(cuda_dev):
    """BASELINE configs[1] at the real width: RandomSearch over N = 64 single-image candidates (T = 50), Oracle
    verifier, candidates and step noise injected exactly as the reference run drew them.  All 64 scores within
    1e-3; the selected index is the reference's whenever its margin exceeds the tolerance, and the pairwise
    ordering of every two candidates whose reference scores differ by more than 2e-3 is reproduced."""
    cfg = cases.LONG_CASES["u_A_search64"]
    g = golden("search_u_A_search64")
    net, _ = build_shell(cfg, cuda_dev)
    smp = _sampler(cfg, net, cuda_dev)
    from its_b200.search import search_algorithm as S
    from its_b200.search import verifier as V
    ver = V.OracleVerifier()
    den = S.make_denoise_fn(smp, None, step_noise=cases.search_noise(cfg).to(cuda_dev), seed=1)
    cands = cases.search_candidates(cfg).to(cuda_dev)
    rs = S.RandomSearch(n_candidates=cfg["n_candidates"])
    best_noise, best_score = rs.search(tuple(cfg["noise_shape"]), den, ver.score, device="cuda", verbose=False,
                                       candidate_noise=cands)
    got, ref = rs.last_scores.cpu().numpy().astype(np.float64), g["rs_scores"]
    print("64-candidate search: max score err %.2e, selected %d (reference %d)" % (np.abs(got - ref).max(),
                                                                                    rs.last_index, int(g["rs_best_index"])))
    assert rs.nfes == 64
    assert np.abs(got - ref).max() <= SCORE_TOL
    order = np.sort(ref)[::-1]
    if order[0] - order[1] > 2 * SCORE_TOL:
        assert rs.last_index == int(g["rs_best_index"])
    # the winner is within tolerance of the reference's best, and it is the noise tensor handed in
    assert ref[rs.last_index] >= order[0] - 2 * SCORE_TOL
    assert torch.equal(best_noise, cands[rs.last_index]) and abs(best_score - ref[rs.last_index]) <= SCORE_TOL
    far = np.abs(ref[:, None] - ref[None, :]) > 2 * SCORE_TOL
    assert np.array_equal((got[:, None] > got[None, :])[far], (ref[:, None] > ref[None, :])[far])


@pytest.mark.parametrize("name", ["u_search", "c_search"])
def test_path_search_restart_vs_oracle(cuda_dev, name):
    """restart=True (BASELINE config 4, "restart at intermediate t") against the oracle: pivot trajectory down to
    injection_step, perturbation of THAT state, remaining steps for every path — injected step noise and
    variations, scores within 1e-3, same selection when the margin allows, same perturbed state returned."""
    cfg = cases.SEARCH_CASES[name]
    dev = cuda_dev
    net, sd = build_shell(cfg, dev)
    smp = _sampler(cfg, net, dev)
    from its_b200.search import search_algorithm as S
    from its_b200.search import verifier as V
    ver = {"oracle": V.OracleVerifier(), "aesthetic": V.AestheticPredictor()}[cfg["verifier"]]
    labels = cases.search_labels(cfg)
    step_noise = cases.search_noise(cfg)
    den = S.make_denoise_fn(smp, None if labels is None else labels.to(dev), step_noise=step_noise.to(dev), seed=1)
    shape = tuple(cfg["noise_shape"])
    rng = np.random.default_rng(9100 + cfg["T"])
    initial = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
    n_paths, inj, scale = 4, cfg["T"] // 2, 0.1
    var = torch.from_numpy(rng.standard_normal((n_paths,) + shape).astype(np.float32))
    ps = S.PathSearch(n_paths=n_paths, injection_step=inj, noise_scale=scale)
    kw = {} if labels is None else {"labels": labels.to(dev)}
    pn, pscore, ph = ps.search(initial.to(dev), den, ver.score, timesteps=cfg["T"], device="cuda",
                               variations=var.to(dev), restart=True, **kw)
    sched = O.schedule(cfg["beta_1"], cfg["beta_T"], cfg["T"])
    ref_noise, ref_score, ref_h = O.path_search_restart(
        sd, sched, initial, list(var), scale, inj, lambda ts: step_noise[ts], O.VERIFIERS[cfg["verifier"]],
        labels=labels, w=cfg.get("w", 0.0))
    got, ref = np.array(ph["scores"]), np.array(ref_h["scores"])
    print("%s restart path search: scores %s vs oracle %s" % (name, got, ref))
    assert np.abs(got - ref).max() <= SCORE_TOL
    assert ph["injection_points"] == ref_h["injection_points"] and ps.nfes == n_paths
    order = np.sort(ref)[::-1]
    if order[0] - order[1] > 2 * SCORE_TOL:
        assert abs(pscore - ref_score) <= SCORE_TOL
        assert rel_err(pn.cpu(), ref_noise) < 5e-3          # the perturbed intermediate state x_inj + 0.1 v


# ---------------------------------------------------------------------------------------------- fp32 mode --
FP32_SAMPLE_TOL = 1e-4      # north_star: "samples within max-abs ... 1e-4 in fp32"


@pytest.mark.parametrize("name", ["u_small", "u_3lvl", "c_small", "u_A", "c_C", "u_E"])
def test_fp32_mode_unet_forward_vs_reference(cuda_dev, name):
    """precision = "fp32" (CUDA-core fp32 kernels, csrc/fp32_path.cu): eps of every forward fixture — small nets and
    the real widths of configs A, C, E — against the reference's fp32 output."""
    cfg = cases.FORWARD_CASES[name]
    net, _ = build_shell(cfg, cuda_dev)
    net.precision = "fp32"
    x, t, labels = cases.forward_inputs(cfg)
    args = (x.to(cuda_dev), t.to(cuda_dev)) + ((labels.to(cuda_dev),) if labels is not None else ())
    eps = net(*args).cpu()
    ref = torch.from_numpy(golden("fwd_" + name)["eps"])
    print("fp32 mode %s: |eps|max %.3e, max abs err %.2e, rel %.2e" % (name, ref.abs().max().item(),
                                                                      (eps - ref).abs().max().item(), rel_err(eps, ref)))
    assert rel_err(eps, ref) < 2e-5
    from its_b200.engine_f32 import UNetPlanF32
    assert all(isinstance(p, UNetPlanF32) for p in net._plans.values())


@pytest.mark.parametrize("name", ["u_small_T20", "c_small_T20"])
def test_fp32_mode_sampler_within_1e4_of_reference(cuda_dev, name):
    """The two sampler fixtures (T = 20, injected noise, guidance w = 1.8 on the conditional one) in fp32 mode:
    every sample within 1e-4 of the reference's, graph replay included."""
    cfg = cases.SAMPLER_CASES[name]
    net, _ = build_shell(cfg, cuda_dev)
    net.precision = "fp32"
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, labels = cases.sampler_inputs(cfg)
    a = (x_T.to(cuda_dev),) + ((labels.to(cuda_dev),) if labels is not None else ())
    x0 = smp(*a, noise=noise.to(cuda_dev)).cpu()
    ref = torch.from_numpy(golden("smp_" + name)["x0"])
    err = (x0 - ref).abs().max().item()
    print("fp32 mode %s: max abs err of the samples %.2e" % (name, err))
    assert err <= FP32_SAMPLE_TOL
    smp.use_cuda_graph = False
    assert torch.equal(smp(*a, noise=noise.to(cuda_dev)).cpu(), x0)      # eager == graph replay, bit for bit


def test_fp32_mode_full_length_trajectory_no_excused_pixels(cuda_dev):
    """The T = 1000 config-A fixture (synthetic O(1) weights; the state grows to |x| ~ 1e3 before the clip) in fp32
    mode: the un-clipped state within 2e-5 (relative) of the reference's at every checkpoint and EVERY one of the
    6144 clipped samples within 2e-2 — the carve-out of the 16-bit test (pixels inside the arithmetic noise of the
    clip boundary) is not needed here."""
    cfg = cases.LONG_CASES["u_A_T1000"]
    g = golden("smp_u_A_T1000")
    net, _ = build_shell(cfg, cuda_dev)
    net.precision = "fp32"
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, _ = cases.sampler_inputs(cfg)
    x, noise = x_T.to(cuda_dev), noise.to(cuda_dev)
    first = cfg["T"] - 1
    for stop in tuple(cfg["keep_at"]) + (0,):
        x = smp(x, noise=noise, t_start=first, t_stop=stop, clip=False)
        ref = torch.from_numpy(g["x0_preclip"] if stop == 0 else g[f"x_after_{stop}"]).to(cuda_dev)
        print("fp32 mode, config A after step %d: |x|max %.3e rel err %.2e" % (stop, ref.abs().max().item(), rel_err(x, ref)))
        assert rel_err(x, ref) < 2e-5, (stop, rel_err(x, ref))
        first = stop - 1
    d = (torch.clip(x, -1, 1) - torch.from_numpy(g["x0"]).to(cuda_dev)).abs()
    print("fp32 mode, config A T=1000: max abs err over all %d clipped samples %.2e" % (d.numel(), d.max().item()))
    assert d.max().item() <= SAMPLE_TOL


@pytest.mark.parametrize("name,fixture", [("c_C_T1000", "smp_c_C_T1000"), ("u_E_T2000", "smp_u_E_T2000")])
def test_fp32_mode_first_segment_of_configs_C_and_E(cuda_dev, name, fixture):
    """fp32 mode on the guided net (CFG w = 1.8, config C's real width) and on the 64x64 net (config E): the first
    segment of the full-length fixtures (steps T-1 .. first checkpoint: 100 resp. 200 steps — the CUDA-core path is
    slow) reproduces the reference's un-clipped state to 1e-5 relative."""
    cfg = cases.LONG_CASES[name]
    g = golden(fixture)
    net, _ = build_shell(cfg, cuda_dev)
    net.precision = "fp32"
    smp = _sampler(cfg, net, cuda_dev)
    x_T, noise, labels = cases.sampler_inputs(cfg)
    stop = cfg["keep_at"][0]
    kw = dict(noise=noise.to(cuda_dev), t_start=cfg["T"] - 1, t_stop=stop, clip=False)
    x = smp(x_T.to(cuda_dev), labels.to(cuda_dev), **kw) if labels is not None else smp(x_T.to(cuda_dev), **kw)
    ref = torch.from_numpy(g[f"x_after_{stop}"]).to(cuda_dev)
    print("fp32 mode %s after step %d: |x|max %.3e rel err %.2e max abs %.2e" % (name, stop, ref.abs().max().item(),
                                                                              rel_err(x, ref), (x - ref).abs().max().item()))
    assert rel_err(x, ref) < 1e-5


def test_fp32_mode_cross_checks_the_tensor_core_plan_on_device(cuda_dev):
    """The two plans evaluate the same shell on the same inputs: the 16-bit tcgen05 plan stays within its error
    budget of the fp32 plan at config A's real width and a 64-image batch (no CPU fixture involved)."""
    cfg = cases.FORWARD_CASES["u_A"]
    net, _ = build_shell(cfg, cuda_dev)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(16, 3, 32, 32, generator=gen).to(cuda_dev)
    t = torch.randint(0, cfg["T"], (16,), generator=gen).to(cuda_dev)
    fast = net(x, t)
    net.precision = "fp32"
    slow = net(x, t)
    print("16-bit plan vs fp32 plan, config A, 16 images: rel err %.2e rms %.2e" % (rel_err(fast, slow), rms_err(fast, slow)))
    assert rms_err(fast, slow) < 1e-2 and rel_err(fast, slow) < 3e-2


# ------------------------------------------------------------------------------ driver / search housekeeping --
def test_fp16_residual_stream_overflow_raises_nan_in_tensor(cuda_dev):
    """The opt-in fp16 residual stream cannot hold a state beyond 65504: the overflow must surface as the sampler's
    own AssertionError("nan in tensor."), never as silently wrong samples; the default bf16 stream runs the same
    input through."""
    from its_b200.Diffusion import GaussianDiffusionSampler
    cfg = cases.SAMPLER_CASES["u_small_T20"]
    x_T, noise, _ = cases.sampler_inputs(cfg)
    big = (x_T * 3e5).to(cuda_dev)
    net, _ = build_shell(cfg, cuda_dev)
    smp = GaussianDiffusionSampler(net, cfg["beta_1"], cfg["beta_T"], cfg["T"]).to(cuda_dev)
    smp.print_steps = False
    out = smp(big, noise=noise.to(cuda_dev))
    assert torch.isfinite(out).all()
    net.residual_fp16 = True
    net.invalidate_plans()
    with pytest.raises(AssertionError, match="nan in tensor"):
        smp(big, noise=noise.to(cuda_dev))


def test_driver_residual_fp16_auto_policy(tmp_path, cuda_dev):
    """`residual_fp16: auto` turns the fp16 residual stream on for a loaded checkpoint at T <= 1000 only; T = 2000
    (an untrained or mismatched net reaches |x_t| ~ 2e5 there) keeps bf16 unless the config forces it."""
    from its_b200 import inference as I
    cfg = dict(cases.SAMPLER_CASES["u_small_T20"])
    net, sd = build_shell(cfg, None)
    ck = tmp_path / "ckpt.pt"
    torch.save(sd, ck)
    base = dict(T=cfg["T"], channel=cfg["ch"], channel_mult=cfg["ch_mult"], attn=cfg["attn"],
                num_res_blocks=cfg["num_res_blocks"], dropout=cfg["dropout"], img_size=cfg["img"])
    for T, mode, path, want in ((1000, "auto", str(ck), True), (2000, "auto", str(ck), False),
                                (2000, True, str(ck), True), (1000, False, str(ck), False), (1000, "auto", None, False)):
        m = I.create_and_load_model(dict(base, T=T, residual_fp16=mode, checkpoint_path=path), cuda_dev)
        assert bool(m.residual_fp16) is want, (T, mode, path)
        assert m.precision == "16bit"
    assert I.create_and_load_model(dict(base, precision="fp32", checkpoint_path=None), cuda_dev).precision == "fp32"


def test_callable_mode_candidates_use_their_own_noise_streams(cuda_dev):
    """Serial (callable) mode: the i-th candidate of a search runs on Philox stream i — fresh step noise per
    candidate like the reference's randn_like (Diffusion.py:96) — i.e. exactly the trajectory population mode
    gives global candidate i, so both modes return the same scores and the same winner."""
    cfg = cases.SEARCH_CASES["u_search"]
    net, _ = build_shell(cfg, cuda_dev)
    smp = _sampler(cfg, net, cuda_dev)
    from its_b200.search import search_algorithm as S
    from its_b200.search import verifier as V
    ver = V.OracleVerifier()
    den = S.make_denoise_fn(smp, None, seed=11)
    shape = tuple(cfg["noise_shape"])
    cands = cases._randn(77, (5,) + shape).to(cuda_dev)
    rs = S.RandomSearch(n_candidates=5)
    n1, s1 = rs.search(shape, den, ver.score, device="cuda", verbose=False, candidate_noise=cands)
    pop = [float(x) for x in rs.last_scores.tolist()]
    seen = []
    rs2 = S.RandomSearch(n_candidates=5)
    n2, s2 = rs2.search(shape, den, lambda im, **kw: seen.append(ver.score(im)) or seen[-1], device="cuda",
                        verbose=False, candidate_noise=cands)
    assert seen == pop and s2 == s1 and torch.equal(n1, n2)
    assert len(set(pop)) == 5                      # different streams: no two candidates share a trajectory
    # the images of the winner are the ones that produced its score
    img = den.denoise_candidates(n1.unsqueeze(0), rs.last_index)[0]
    assert ver.score(img) == s1


def test_out_of_range_labels_and_steps_raise_like_nn_embedding(cuda_dev):
    """ModelCondition's tables are nn.Embedding: an index outside them raises IndexError in the reference; the shell
    raises the same error instead of reading a neighbouring row (and the kernel itself marks such a row NaN)."""
    cfg = cases.SAMPLER_CASES["c_small_T20"]
    net, _ = build_shell(cfg, cuda_dev)
    x, t, labels = cases.forward_inputs(cfg)
    x, t, labels = x.to(cuda_dev), t.to(cuda_dev), labels.to(cuda_dev)
    net(x, t, labels)
    with pytest.raises(IndexError):
        net(x, t, labels + cfg["num_labels"] + 1)
    with pytest.raises(IndexError):
        net(x, t + cfg["T"], labels)
    smp = _sampler(cfg, net, cuda_dev)
    with pytest.raises(IndexError):
        smp(x, torch.full_like(labels, cfg["num_labels"] + 3))
