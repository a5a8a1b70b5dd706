"""GPU: every kernel of libits_b200 against a plain torch fp32 statement of the same
op (inputs/weights pre-rounded to bf16 where the kernel stores bf16), called
through the C-ABI."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ddpm_oracle as O
from oracle import philox
from tests import cases

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):  # NHWC any -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def check_close(got, ref, tol, what=""):
    """max |got-ref| <= tol * max|ref| (bf16 storage dominates the error budget)."""
    err = (got - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-6)
    assert err <= tol * scale, f"{what}: max err {err:.4e} vs scale {scale:.4e} (tol {tol})"


# ------------------------------------------------------------ DDPM step ------
def _coef(T, dev):
    from its_b200.Diffusion import GaussianDiffusionSampler
    s = GaussianDiffusionSampler(torch.nn.Identity(), 1e-4, 0.02, T)
    return s, s._coef_table(dev)


@pytest.mark.parametrize("guided", [False, True])
def test_ddpm_step_bit_exact_with_injected_noise(cuda_dev, built_lib, guided):
    from its_b200 import _lib
    T, B, n = 50, 5, 3 * 16 * 16
    smp, coef = _coef(T, cuda_dev)
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(B, n, generator=g).to(cuda_dev)
    e_c = torch.randn(B, n, generator=g).to(cuda_dev)
    e_u = torch.randn(B, n, generator=g).to(cuda_dev) if guided else None
    noise = torch.randn(T, B, n, generator=g).to(cuda_dev)
    w = 1.8
    for t in (T - 1, 7, 1, 0):
        tt = torch.full((B,), t, dtype=torch.long, device=cuda_dev)
        # reference expression order (Diffusion.py:67-72,99; DiffusionCondition.py:85)
        eps = (1. + w) * e_c - w * e_u if guided else e_c
        var = torch.cat([smp.posterior_var[1:2], smp.betas[1:]]).to(cuda_dev)
        c1 = smp.coeff1.to(cuda_dev)[tt].float().view(B, 1)
        c2 = smp.coeff2.to(cuda_dev)[tt].float().view(B, 1)
        mean = c1 * x - c2 * eps
        ref = mean + torch.sqrt(var[tt].float().view(B, 1)) * noise[t] if t > 0 else mean
        ref_clip = torch.clip(ref, -1, 1) if t == 0 else ref
        xk = x.clone()
        t_dev = torch.tensor([t], dtype=torch.int32, device=cuda_dev)
        flag = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
        _lib.check(built_lib.its_ddpm_step(xk.data_ptr(), e_c.data_ptr(), e_u.data_ptr() if guided else None,
                                           noise.data_ptr(), B * n, B, n, coef.data_ptr(), t_dev.data_ptr(), w,
                                           0, 0, flag.data_ptr(), 1, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(xk, ref_clip), f"t={t}: max diff {(xk - ref_clip).abs().max().item()}"
        assert flag.item() == 0


def test_ddpm_step_nan_flag_and_advance(cuda_dev, built_lib):
    from its_b200 import _lib
    _, coef = _coef(10, cuda_dev)
    x = torch.zeros(2, 64, device=cuda_dev)
    e = torch.zeros(2, 64, device=cuda_dev)
    e[1, 5] = float("nan")
    t_dev = torch.tensor([3], dtype=torch.int32, device=cuda_dev)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    _lib.check(built_lib.its_ddpm_step(x.data_ptr(), e.data_ptr(), None, None, 0, 2, 64, coef.data_ptr(),
                                       t_dev.data_ptr(), 0.0, 1, 0, flag.data_ptr(), 1, _lib.stream_ptr()))
    _lib.check(built_lib.its_step_advance(t_dev.data_ptr(), -1, _lib.stream_ptr()))
    assert flag.item() == 1 and t_dev.item() == 2


def test_philox_matches_numpy_restatement(cuda_dev, built_lib):
    from its_b200.search.search_algorithm import philox_normal
    seed, cand0, tag = 0x1234567890ABCDEF & (2 ** 62 - 1), 5, 0x40000000
    z = philox_normal((6, 3, 8, 8), seed, cand0, tag, cuda_dev).cpu().numpy().reshape(6, -1)
    ref = philox.normal(seed, cand0, tag, 6, 192)
    assert np.abs(z - ref).max() < 2e-4          # fast log / sincos intrinsics vs libm
    base = torch.arange(192, dtype=torch.float32, device=cuda_dev)
    z2 = philox_normal((6, 3, 8, 8), seed, cand0, tag, cuda_dev, base=base, scale=0.05).cpu().numpy().reshape(6, -1)
    assert np.abs(z2 - (np.arange(192, dtype=np.float32)[None] + 0.05 * ref)).max() < 1e-4
    big = philox_normal((64, 3, 32, 32), 99, 0, 7, cuda_dev)
    assert abs(big.mean().item()) < 5e-3 and abs(big.std().item() - 1) < 5e-3
    # a unit's stream depends on its global id only
    part = philox_normal((2, 3, 32, 32), 99, 10, 7, cuda_dev)
    assert torch.equal(part, big[10:12])


def test_ddpm_step_philox_noise_is_the_library_stream(cuda_dev, built_lib):
    from its_b200 import _lib
    T, B, n = 30, 3, 256
    _, coef = _coef(T, cuda_dev)
    x = torch.zeros(B, n, device=cuda_dev)
    e = torch.zeros(B, n, device=cuda_dev)
    t = 11
    t_dev = torch.tensor([t], dtype=torch.int32, device=cuda_dev)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    _lib.check(built_lib.its_ddpm_step(x.data_ptr(), e.data_ptr(), None, None, 0, B, n, coef.data_ptr(),
                                       t_dev.data_ptr(), 0.0, 77, 4, flag.data_ptr(), 0, _lib.stream_ptr()))
    ref = coef[t, 2].item() * philox.normal(77, 4, t, B, n)
    assert np.abs(x.cpu().numpy() - ref).max() < 1e-4


# ------------------------------------------------------------ GroupNorm ------
@pytest.mark.parametrize("B,H,C0,C1,silu", [(2, 32, 64, 0, True), (3, 16, 128, 64, True), (8, 4, 512, 512, True),
                                            (2, 8, 384, 0, False), (1, 64, 128, 128, True)])
def test_group_norm(cuda_dev, built_lib, B, H, C0, C1, silu):
    from its_b200.engine import UNetPlan
    g = torch.Generator().manual_seed(B * 1000 + C0 + C1)
    x0 = (torch.randn(B, C0, H, H, generator=g) * 2 + 0.5).to(cuda_dev)
    x1 = (torch.randn(B, C1, H, H, generator=g) * 0.5 - 1).to(cuda_dev) if C1 else None
    gn = torch.nn.GroupNorm(32, C0 + C1).to(cuda_dev)
    with torch.no_grad():
        gn.weight.copy_(1 + 0.2 * torch.randn(C0 + C1, generator=g))
        gn.bias.copy_(0.1 * torch.randn(C0 + C1, generator=g))
    plan = UNetPlan.scratch(cuda_dev, B)
    srcs = [nhwc(x0)] + ([nhwc(x1)] if C1 else [])   # kept alive until after plan.run()
    out = plan.group_norm(srcs, gn, silu)
    plan.run()
    xin = torch.cat([bf(x0)] + ([bf(x1)] if C1 else []), 1)
    with torch.no_grad():
        ref = gn(xin)
        ref = ref * torch.sigmoid(ref) if silu else ref
    check_close(nchw(out), ref, 6e-3, "group_norm")


# ----------------------------------------------------------- head / tail -----
def test_conv_head_and_tail(cuda_dev, built_lib):
    from its_b200 import _lib
    g = torch.Generator().manual_seed(5)
    B, H, ch = 4, 16, 64
    x = torch.randn(2, 3, H, H, generator=g).to(cuda_dev) * 30      # x_t is large at early steps
    w = torch.randn(ch, 3, 3, 3, generator=g).to(cuda_dev) * 0.2
    b = torch.randn(ch, generator=g).to(cuda_dev) * 0.1
    out = torch.empty(B, H, H, ch, dtype=torch.bfloat16, device=cuda_dev)
    _lib.check(built_lib.its_conv_head(out.data_ptr(), x.data_ptr(), w.data_ptr(), b.data_ptr(), B, 2, H, H, 3, ch,
                                       _lib.stream_ptr()))
    ref = F.conv2d(torch.cat([x, x]), w, b, padding=1)
    check_close(nchw(out), ref, 5e-3, "head")
    a = torch.randn(B, ch, H, H, generator=g).to(cuda_dev)
    wt = torch.randn(3, ch, 3, 3, generator=g).to(cuda_dev) * 0.05
    bt = torch.randn(3, generator=g).to(cuda_dev)
    eps = torch.empty(B, 3, H, H, device=cuda_dev)
    _lib.check(built_lib.its_conv_tail(eps.data_ptr(), nhwc(a).data_ptr(), wt.data_ptr(), bt.data_ptr(), B, H, H, ch, 3,
                                       _lib.stream_ptr()))
    check_close(eps, F.conv2d(bf(a), wt, bt, padding=1), 1e-5, "tail")



def test_head_on_tensor_cores_keeps_fp32_grade_accuracy(cuda_dev, built_lib):
    """Head conv (Model.py:269) as hi/lo bf16 patches x hi/lo weights on the tap-GEMM: the only
    rounding left is the bf16 store of the result."""
    from its_b200.engine import UNetPlan
    g = torch.Generator().manual_seed(6)
    B, H, ch = 6, 32, 128
    x = torch.randn(3, 3, H, H, generator=g).to(cuda_dev) * 30      # x_t is large at early steps
    w = torch.randn(ch, 3, 3, 3, generator=g).to(cuda_dev) * 0.2
    b = torch.randn(ch, generator=g).to(cuda_dev) * 0.1
    plan = UNetPlan.scratch(cuda_dev, B)
    out = plan.head_conv(w, b, x, B, H, H)
    assert plan.op_info[0][0] == "head_patches" and out.data_ptr() in plan.stats_of
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(torch.cat([x, x]), w, b, padding=1)
    got = nchw(out)
    assert torch.equal(got, ref.to(torch.bfloat16).float()) or \
        (got - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()
    # a bf16-only GEMM would be ~10x worse than the bf16 store rounding bound checked here
    frac_exact = (got == ref.to(torch.bfloat16).float()).float().mean().item()
    assert frac_exact > 0.98, frac_exact

# -------------------------------------------------------------- tap-GEMM -----
def _conv_case(dev, impl, B, H, Cin, Cout, *, k=3, stride=1, extras=False, seed=0, alpha=0.5):
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(seed + 17 * B + H + Cin + Cout)
    x = torch.randn(B, Cin, H, H, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    plan = UNetPlan.scratch(dev, B, impl)
    kw = {}
    Ho = H // stride
    ref = F.conv2d(bf(x), bf(w), bias, stride=stride, padding=k // 2)
    if extras:
        vec = torch.randn(B, Cout + 8, generator=g).to(dev)
        vec2 = torch.randn(1, Cout, generator=g).to(dev)
        res = torch.randn(B, Cout, Ho, Ho, generator=g).to(dev)
        kw = dict(vec=vec, vec_off=8, vec2=vec2, res=nhwc(res), alpha=alpha)
        ref = alpha * F.conv2d(bf(x), bf(w), None, stride=stride, padding=k // 2) + bias.view(1, -1, 1, 1) \
            + vec[:, 8:].view(B, Cout, 1, 1) + vec2.view(1, Cout, 1, 1) + bf(res)
    wp = pack_conv_weight(w).contiguous()       # fp32: the plan packs it in the sources' 16-bit format
    xin = nhwc(x)
    out = plan.conv([(xin, Cin, 0, stride, False)], [(taps_square(k), 0, 0, 0)], Ho, Ho, wp, Cout, bias=bias, **kw)
    plan.run()
    torch.cuda.synchronize()
    return nchw(out), ref


CONV_SHAPES = [
    # B, H, Cin, Cout, k, stride, extras
    (2, 32, 64, 64, 3, 1, False),      # bn=64, 8 tiles per image
    (2, 32, 128, 128, 3, 1, True),     # bn=128 + full epilogue
    (2, 16, 256, 256, 3, 1, False),    # bn=256
    (4, 8, 128, 192, 3, 1, True),      # bn=192, two images per tile
    (8, 4, 512, 512, 3, 1, False),     # eight images per tile, 72 k-blocks
    (3, 8, 64, 128, 3, 1, False),      # ragged last tile (batch not a multiple of the box)
    (2, 4, 128, 128, 3, 1, False),     # box larger than the batch
    (2, 64, 128, 128, 3, 1, False),    # 64x64 rows: box 64x2
    (2, 16, 128, 384, 1, 1, True),     # 1x1, three N tiles
    (2, 32, 128, 128, 3, 2, False),    # DownSample stride 2 (TMA element strides)
    (4, 16, 256, 256, 3, 2, True),
    (2, 32, 64, 64, 5, 2, False),      # 5x5 stride 2
]


@pytest.mark.parametrize("impl", [1, 0], ids=["cudacore", "tcgen05"])
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,extras", CONV_SHAPES)
def test_conv_igemm_vs_torch(cuda_dev, built_lib, impl, B, H, Cin, Cout, k, stride, extras):
    got, ref = _conv_case(cuda_dev, impl, B, H, Cin, Cout, k=k, stride=stride, extras=extras)
    check_close(got, ref, 6e-3, f"conv impl={impl}")



@pytest.mark.parametrize("schedule", [1, 2], ids=["tile_per_cta", "persistent"])
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,extras", CONV_SHAPES)
def test_conv_igemm_schedules(cuda_dev, built_lib, schedule, B, H, Cin, Cout, k, stride, extras):
    """Both tcgen05 schedules on every shape (the persistent one folds the residual as an identity tap)."""
    from its_b200.engine import UNetPlan
    orig = UNetPlan.scratch

    def scratch(dev, n, impl=None):
        p = orig(dev, n, impl)
        p.schedule = schedule
        p.split_k = schedule != 2
        return p
    UNetPlan.scratch = scratch
    try:
        got, ref = _conv_case(cuda_dev, 0, B, H, Cin, Cout, k=k, stride=stride, extras=extras,
                             alpha=1.0 if schedule == 2 else 0.5)
    finally:
        UNetPlan.scratch = orig
    check_close(got, ref, 6e-3, f"conv schedule={schedule}")


@pytest.mark.parametrize("B,H,C0,C1,Cout", [(3, 32, 128, 0, 128), (2, 16, 256, 128, 256), (5, 8, 384, 256, 384),
                                            (9, 4, 512, 512, 512), (2, 64, 128, 0, 128)])
def test_persistent_conv_statistics_feed_group_norm(cuda_dev, built_lib, B, H, C0, C1, Cout):
    """The persistent tap-GEMM leaves per-(image, 4 channels, tile) sums of the stored bf16 tensor;
    GroupNorm of (that tensor | a second such tensor) through its_group_norm_apply must match
    nn.GroupNorm on the stored values (Model.py:170-173 after Model.py:279-280's concat)."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(B + H + C0 + C1)
    plan = UNetPlan.scratch(cuda_dev, B, 0)
    plan.split_k = False
    plan.schedule = 2
    keep, outs = [], []
    for C in [c for c in (C0, C1) if c]:
        x = (torch.randn(B, 64, H, H, generator=g) * 1.5 + 0.3).to(cuda_dev)
        w = (torch.randn(C, 64, 3, 3, generator=g) / 24).to(cuda_dev)
        bias = torch.randn(C, generator=g).to(cuda_dev)
        xin, wp = nhwc(x), pack_conv_weight(w).to(torch.bfloat16).contiguous()
        keep += [xin, wp, bias]
        outs.append(plan.conv([(xin, 64, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, wp, C, bias=bias))
    gn = torch.nn.GroupNorm(32, C0 + C1).to(cuda_dev)
    with torch.no_grad():
        gn.weight.copy_(1 + 0.2 * torch.randn(C0 + C1, generator=g))
        gn.bias.copy_(0.1 * torch.randn(C0 + C1, generator=g))
    y = plan.group_norm(outs, gn, True)
    assert plan.op_info[-1][0] == "group_norm_apply"
    plan.run()
    torch.cuda.synchronize()
    # the partial sums are sums of the stored values
    for o in outs:
        st, parts = plan.stats_of[o.data_ptr()]
        v = o.float()
        C = v.shape[-1]
        ref_s = v.reshape(B, -1, C // 4, 4).sum((1, 3))
        ref_q = (v * v).reshape(B, -1, C // 4, 4).sum((1, 3))
        assert (st[..., 0].sum(1) - ref_s).abs().max().item() <= 2e-3 * max(1.0, ref_s.abs().max().item())
        assert (st[..., 1].sum(1) - ref_q).abs().max().item() <= 2e-4 * ref_q.abs().max().item()
    with torch.no_grad():
        ref = gn(torch.cat([nchw(o) for o in outs], 1))
        ref = ref * torch.sigmoid(ref)
    check_close(nchw(y), ref, 6e-3, "group_norm_apply")

FUSED_GN_SHAPES = [
    # B, H, Cin, Cout, silu, gn_only, extras
    (8, 4, 512, 512, True, True, False),      # 4x4: eight whole images per tile, split-K owner re-reads its partial
    (19, 4, 256, 512, True, False, True),     # ragged last tile, raw + normalised output
    (4, 8, 128, 384, True, True, True),       # 8x8: two images per tile, 12-channel groups inside a 192-wide N tile
    (5, 8, 384, 256, False, False, False),    # no Swish (the attention block's GroupNorm)
    (3, 16, 128, 256, True, True, True),      # 16x16: an image = 2 tiles, peer rendezvous
    (2, 16, 64, 128, True, False, False),     # 16x16, 128 channels: two row boxes per tile (MT = 2, transposed accumulator)
    (3, 32, 128, 128, True, True, True),      # 32x32 transposed-accumulator mode: 4 peer tiles per image
    (50, 32, 64, 128, True, False, True),     # 200 tiles on 148 CTAs: peers straddle the round boundary
    (1, 64, 64, 128, True, True, False),      # 64x64: 16 peer tiles
    (2, 32, 128, 256, False, False, False),   # 32x32 with a 256-wide tile: 8 peers
]


@pytest.mark.parametrize("B,H,Cin,Cout,silu,gn_only,extras", FUSED_GN_SHAPES)
def test_conv_with_fused_group_norm_epilogue(cuda_dev, built_lib, B, H, Cin, Cout, silu, gn_only, extras):
    """GroupNorm(32, Cout) (+ Swish) of a convolution's output applied by that launch's own epilogue
    (its_conv_desc.gn_out; Model.py:186-190 reads exactly what Model.py:173 wrote) against nn.GroupNorm on the
    fp32 convolution: whole-image tiles, peer-tile rendezvous, transposed accumulators, split-K, raw + normalised
    or normalised only.  The launch must replace the stand-alone GroupNorm launch, and repeated launches (graph
    replay) must find the rendezvous counters zero again."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    dev = cuda_dev
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cout)
    x = (torch.randn(B, Cin, H, H, generator=g) * 1.3 + 0.2).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    gn = torch.nn.GroupNorm(32, Cout).to(dev)
    with torch.no_grad():
        gn.weight.copy_(1 + 0.3 * torch.randn(Cout, generator=g))
        gn.bias.copy_(0.2 * torch.randn(Cout, generator=g))
    plan = UNetPlan.scratch(dev, B, 0)
    kw = {}
    ref = F.conv2d(bf(x), bf(w), bias, padding=1)
    if extras:
        vec = torch.randn(B, Cout + 8, generator=g).to(dev)
        vec2 = torch.randn(1, Cout, generator=g).to(dev)
        kw = dict(vec=vec, vec_off=8, vec2=vec2)
        ref = ref + vec[:, 8:].view(B, Cout, 1, 1) + vec2.view(1, Cout, 1, 1)
    wp = pack_conv_weight(w).contiguous()
    xin = nhwc(x)                          # the plan borrows the pointer: keep the tensor alive
    out = plan.conv([(xin, Cin, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, wp, Cout, bias=bias,
                    fuse_gn=(gn, silu), gn_only=gn_only, **kw)
    assert plan.n_gn_fused == 1, "the layer was expected to be fusable"
    y = plan.group_norm([out], gn, silu)
    assert len(plan.ops) == 1 and plan.op_info[-1][0] == "tapgemm_sm100"     # no GroupNorm launch was appended
    with torch.no_grad():
        want = gn(ref)
        if silu:
            want = want * torch.sigmoid(want)
    for rep in range(3):                   # repeated launches: the rendezvous counters keep counting
        y.zero_()
        plan.run()
        torch.cuda.synchronize()
        got = nchw(y)
        bad = ~torch.isfinite(got) | ((got - want).abs() > 8e-3 * want.abs().max())
        if bad.any():          # say where: images / 64-channel tiles that are wrong
            imgs = sorted(set(bad.nonzero()[:, 0].tolist()))
            tiles = sorted(set((bad.nonzero()[:, 1] // 64).tolist()))
            print(f"launch {rep}: {int(bad.sum())} bad of {bad.numel()}, nan {int((~torch.isfinite(got)).sum())}, "
                  f"zeros among bad {int((got[bad] == 0).sum())}, images {imgs[:8]}, channel tiles {tiles}")
        check_close(got, want, 8e-3, f"fused GroupNorm output, launch {rep}")
    assert y.dtype == torch.float16
    if gn_only:
        assert out is y                    # the raw tensor is not materialised
    else:
        check_close(nchw(out), ref, 6e-3, "raw output")
        # the statistics by-product still serves later consumers of the raw tensor (skip connections)
        st, parts = plan.stats_of[out.data_ptr()]
        v = out.float()
        ref_s = v.reshape(B, -1, Cout // 4, 4).sum((1, 3))
        assert (st[..., 0].sum(1) - ref_s).abs().max().item() <= 2e-3 * max(1.0, ref_s.abs().max().item())
    for t in plan.keep:                    # every rendezvous counter saw all its peer tiles in each of the 3 launches
        if t.dtype == torch.int32:
            assert int(t.min().item()) == int(t.max().item()) and int(t[0].item()) % 3 == 0 and int(t[0].item()) > 0


def test_fused_group_norm_matches_the_two_launch_plan(cuda_dev, built_lib):
    """The same ResBlock-shaped chain built with and without the fused epilogue (ITS_GN_FUSION): the second conv's
    output agrees to the rounding of the 16-bit intermediate the fusion removes."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    dev = cuda_dev
    B, H, Cc = 6, 16, 256
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, Cc, H, H, generator=g).to(dev)
    w1 = (torch.randn(Cc, Cc, 3, 3, generator=g) / math.sqrt(Cc * 9)).to(dev)
    w2 = (torch.randn(Cc, Cc, 3, 3, generator=g) / math.sqrt(Cc * 9)).to(dev)
    gn = torch.nn.GroupNorm(32, Cc).to(dev)
    outs, xin = {}, nhwc(x)
    for fused in (False, True):
        plan = UNetPlan.scratch(dev, B, 0)
        plan.gn_fusion = "all" if fused else "off"
        h1 = plan.conv([(xin, Cc, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, pack_conv_weight(w1).contiguous(),
                       Cc, fuse_gn=(gn, True), gn_only=True)
        a2 = plan.group_norm([h1], gn, True)
        h2 = plan.conv([(a2, Cc, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, pack_conv_weight(w2).contiguous(), Cc)
        assert len(plan.ops) == (2 if fused else 3)
        plan.run()
        torch.cuda.synchronize()
        outs[fused] = nchw(h2)
    check_close(outs[True], outs[False], 6e-3, "fused vs two-launch chain")


@pytest.mark.parametrize("impl", [1, 0], ids=["cudacore", "tcgen05"])
def test_conv_fused_shortcut_three_sources(cuda_dev, built_lib, impl):
    """ResBlock conv2 + 1x1 shortcut over the (h | skip) concat as extra K (Model.py:191-207)."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(9)
    B, H, C, Ch, Cs = 2, 16, 128, 128, 64
    a2 = torch.randn(B, C, H, H, generator=g).to(cuda_dev)
    h = torch.randn(B, Ch, H, H, generator=g).to(cuda_dev)
    sk = torch.randn(B, Cs, H, H, generator=g).to(cuda_dev)
    w2 = (torch.randn(C, C, 3, 3, generator=g) / 34).to(cuda_dev)
    ws = (torch.randn(C, Ch + Cs, 1, 1, generator=g) / 14).to(cuda_dev)
    bias = torch.randn(C, generator=g).to(cuda_dev)
    plan = UNetPlan.scratch(cuda_dev, B, impl)
    wp = torch.cat([pack_conv_weight(w2), ws[:, :, 0, 0]], 1).to(torch.bfloat16).contiguous()
    taps = taps_square(3) + [(1, 0, 0), (2, 0, 0)]
    ins = [nhwc(a2), nhwc(h), nhwc(sk)]
    out = plan.conv([(ins[0], C, 0, 1, False), (ins[1], Ch, 0, 1, False), (ins[2], Cs, 0, 1, False)],
                    [(taps, 0, 0, 0)], H, H, wp, C, bias=bias)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(bf(a2), bf(w2), bias, padding=1) + F.conv2d(torch.cat([bf(h), bf(sk)], 1), bf(ws))
    check_close(nchw(out), ref, 6e-3, "fused shortcut")


@pytest.mark.parametrize("impl", [1, 0], ids=["cudacore", "tcgen05"])
def test_conv_tail_on_tensor_cores_nchw_output(cuda_dev, built_lib, impl):
    """3-channel tail conv (Model.py:257-262) through the tap-GEMM with the NCHW fp32 epilogue."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(21)
    B, H, C = 3, 32, 128
    a = torch.randn(B, C, H, H, generator=g).to(cuda_dev)
    w = (torch.randn(3, C, 3, 3, generator=g) / 34).to(cuda_dev)
    bias = torch.randn(3, generator=g).to(cuda_dev)
    plan = UNetPlan.scratch(cuda_dev, B, impl)
    ain = nhwc(a)
    out = torch.full((B, 3, H, H), float("nan"), device=cuda_dev)
    plan.conv([(ain, C, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H,
              pack_conv_weight(w).to(torch.bfloat16).contiguous(), 3, bias=bias, out=out, out_fp32=True, out_nchw=True)
    plan.run()
    torch.cuda.synchronize()
    check_close(out, F.conv2d(bf(a), bf(w), bias, padding=1), 2e-3, "tail")


def test_split_k_matches_unsplit_bitwise_inputs(cuda_dev, built_lib):
    """Split-K (few output tiles) against the un-split launch of the same layer."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(4)
    B, H, C = 16, 4, 512      # 4x4 maps: 16 CTAs at the design batch -> split 9
    x = torch.randn(B, C, H, H, generator=g).to(cuda_dev)
    w = (torch.randn(C, C, 3, 3, generator=g) / 68).to(cuda_dev)
    xin, wp = nhwc(x), pack_conv_weight(w).to(torch.bfloat16).contiguous()
    outs = []
    for split in (True, False):
        plan = UNetPlan.scratch(cuda_dev, B, 0)
        plan.split_k = split
        outs.append(plan.conv([(xin, C, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, wp, C))
        assert (plan.descs[0].splits > 1) == split
        plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(bf(x), bf(w), None, padding=1)
    check_close(nchw(outs[0]), ref, 6e-3, "split")
    check_close(nchw(outs[0]), nchw(outs[1]), 8e-3, "split vs unsplit")   # fp32 re-association + bf16 rounding


def test_persistent_split_k_with_more_work_items_than_sms(cuda_dev, built_lib):
    """Persistent split-K at a batch whose work items (tiles x splits) exceed the resident CTAs: partial
    items precede owner items on every CTA, so owners never starve (conv_persist_sm100.cu); the
    result is bit-identical to the same images evaluated in a small batch (fixed summation order)."""
    from its_b200.engine import UNetPlan, pack_conv_weight, taps_square
    g = torch.Generator().manual_seed(9)
    B, H, C = 160, 4, 512
    x = torch.randn(B, C, H, H, generator=g).to(cuda_dev)
    w = (torch.randn(C, C, 3, 3, generator=g) / 68).to(cuda_dev)
    xin, wp = nhwc(x), pack_conv_weight(w).contiguous()
    plan = UNetPlan.scratch(cuda_dev, B, 0)
    out = plan.conv([(xin, C, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, wp, C)
    d = plan.descs[0]
    assert d.splits > 1 and d.schedule != 1 and d.stats_parts > 0     # persistent, split, statistics kept
    assert (B // 8) * (C // d.bn) * d.splits > 148
    for _ in range(3):            # flags are self-cleaning: repeated launches give the same answer
        plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(bf(x), bf(w), None, padding=1)
    check_close(nchw(out), ref, 6e-3, "large-batch split")
    small = UNetPlan.scratch(cuda_dev, 8, 0)
    xs = xin[:8].contiguous()
    out8 = small.conv([(xs, C, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, wp, C)
    small.run()
    torch.cuda.synchronize()
    assert small.descs[0].splits == d.splits
    assert torch.equal(out8, out[:8])


class _Holder:
    pass


@pytest.mark.parametrize("impl", [1, 0], ids=["cudacore", "tcgen05"])
@pytest.mark.parametrize("kind", ["down_uncond", "down_cond", "up_uncond", "up_cond"])
def test_resample_blocks(cuda_dev, built_lib, impl, kind):
    """DownSample / UpSample of both networks through the engine's own block builders."""
    from its_b200.engine import UNetPlan
    torch.manual_seed(3)
    B, H, C = 2, 16, 128
    x = torch.randn(B, C, H, H, device=cuda_dev)
    plan = UNetPlan.scratch(cuda_dev, B, impl)
    m = _Holder()
    xb = bf(x)
    xin = nhwc(x)            # the plan stores raw pointers: inputs must outlive plan.run()
    with torch.no_grad():
        if kind == "down_uncond":
            m.main = torch.nn.Conv2d(C, C, 3, 2, 1).to(cuda_dev)
            out = plan._down(m, xin)
            ref = F.conv2d(xb, bf(m.main.weight), m.main.bias, stride=2, padding=1)
        elif kind == "down_cond":
            m.c1 = torch.nn.Conv2d(C, C, 3, 2, 1).to(cuda_dev)
            m.c2 = torch.nn.Conv2d(C, C, 5, 2, 2).to(cuda_dev)
            out = plan._down(m, xin)
            ref = F.conv2d(xb, bf(m.c1.weight), m.c1.bias, stride=2, padding=1) + \
                F.conv2d(xb, bf(m.c2.weight), m.c2.bias, stride=2, padding=2)
        elif kind == "up_uncond":
            m.main = torch.nn.Conv2d(C, C, 3, 1, 1).to(cuda_dev)
            out = plan._up(m, xin)
            ref = F.conv2d(F.interpolate(xb, scale_factor=2, mode="nearest"), m.main.weight, m.main.bias, padding=1)
        else:
            m.c = torch.nn.Conv2d(C, C, 3, 1, 1).to(cuda_dev)
            m.t = torch.nn.ConvTranspose2d(C, C, 5, 2, 2, 1).to(cuda_dev)
            out = plan._up(m, xin)
            y = F.conv_transpose2d(xb, bf(m.t.weight), m.t.bias, stride=2, padding=2, output_padding=1)
            ref = F.conv2d(bf(y), bf(m.c.weight), m.c.bias, padding=1)
        plan.run()
    tol = 1.2e-2 if kind == "up_uncond" else 8e-3   # folded weights are summed before bf16 rounding
    check_close(nchw(out), ref, tol, kind)


@pytest.mark.parametrize("impl", [1, 0], ids=["cudacore", "tcgen05"])
@pytest.mark.parametrize("B,H,C", [(2, 16, 128), (2, 16, 64), (8, 8, 128), (8, 4, 512), (1, 32, 128), (3, 32, 64),
                                   (2, 8, 384), (5, 32, 128), (2, 16, 384), (3, 16, 256), (3, 4, 512), (5, 8, 256),
                                   (20, 4, 64)])
def test_attention_block(cuda_dev, built_lib, impl, B, H, C):
    """AttnBlock (Model.py:145-164): tensor-core batched GEMM path for >= 128 tokens,
    one-kernel path for small maps."""
    from its_b200.engine import UNetPlan
    torch.manual_seed(B * 100 + H + C)
    at = _Holder()
    at.group_norm = torch.nn.GroupNorm(32, C).to(cuda_dev)
    for nme in ("proj_q", "proj_k", "proj_v", "proj"):
        conv = torch.nn.Conv2d(C, C, 1).to(cuda_dev)
        with torch.no_grad():
            conv.weight.mul_(3.0)
        setattr(at, nme, conv)
    x = torch.randn(B, C, H, H, device=cuda_dev)
    plan = UNetPlan.scratch(cuda_dev, B, impl)
    xin = nhwc(x)
    with torch.no_grad():
        out = plan._attn_block(at, xin)
        plan.run()
        sd = {"a.group_norm.weight": at.group_norm.weight, "a.group_norm.bias": at.group_norm.bias}
        for nme in ("proj_q", "proj_k", "proj_v", "proj"):
            sd[f"a.{nme}.weight"] = getattr(at, nme).weight
            sd[f"a.{nme}.bias"] = getattr(at, nme).bias
        ref = O._attn(sd, "a", bf(x), O._Q(None))
    check_close(nchw(out), ref, 1.5e-2, "attention")


@pytest.mark.parametrize("B,H,C", [(2, 16, 128), (2, 16, 384), (2, 32, 128)])
def test_attention_v_transposed_operand_path_matches_mn_major_path(cuda_dev, built_lib, B, H, C, monkeypatch):
    """The fused cores read V either MN-major from the q|k|v tensor (default) or K-major from a V^T
    projection (ITS_ATTN_VT=1): same products, same order -> the block outputs agree to bf16 rounding of
    one intermediate (V is rounded once either way, the projections are separate launches)."""
    from its_b200.engine import UNetPlan
    torch.manual_seed(7)
    at = _Holder()
    at.group_norm = torch.nn.GroupNorm(32, C).to(cuda_dev)
    for nme in ("proj_q", "proj_k", "proj_v", "proj"):
        setattr(at, nme, torch.nn.Conv2d(C, C, 1).to(cuda_dev))
    x = torch.randn(B, C, H, H, device=cuda_dev)
    xin = nhwc(x)
    outs = []
    for vt in ("0", "1"):
        monkeypatch.setenv("ITS_ATTN_VT", vt)
        plan = UNetPlan.scratch(cuda_dev, B, 0)
        with torch.no_grad():
            out = plan._attn_block(at, xin)
            plan.run()
        kinds = [k for k, _, _ in plan.op_info]
        assert ("attention_fused" in kinds) or ("attention_flash" in kinds)
        assert len(kinds) == (4 if vt == "0" else 5)        # GN, qkv, core, proj (+ V^T projection)
        outs.append(nchw(out).clone())
    torch.cuda.synchronize()
    check_close(outs[0], outs[1], 4e-3, "mn-major vs V^T")


@pytest.mark.parametrize("n_img,N,C", [(8, 16, 512), (3, 16, 64), (19, 16, 128), (5, 32, 256), (4, 64, 384), (9, 64, 512)])
def test_attention_group_kernel_vs_torch(cuda_dev, built_lib, n_img, N, C):
    """its_attention_group directly (the engine uses it for 8x8 maps only): 128 / N images per tile with
    a block-diagonal softmax mask, ragged last group included."""
    from its_b200 import _lib
    g = torch.Generator().manual_seed(n_img * 1000 + N + C)
    qkv = (torch.randn(n_img, N, 3 * C, generator=g) * 1.5).to(cuda_dev).to(torch.bfloat16)
    bias = torch.randn(C, generator=g).to(cuda_dev)
    out = torch.full((n_img, N, C), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    scale = float(C) ** -0.5
    _lib.check(built_lib.its_attention_group(out.data_ptr(), qkv.data_ptr(), bias.data_ptr(), n_img, N, C, scale,
                                             _lib.stream_ptr()), "its_attention_group")
    q, k, v = (t.float() for t in qkv.split(C, dim=-1))
    p = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * scale, dim=-1)
    ref = torch.bmm(p.to(torch.bfloat16).float(), v) + bias
    check_close(out.float(), ref, 1.2e-2, "attention_group")


def test_softmax_rows(cuda_dev, built_lib):
    from its_b200 import _lib
    s = torch.randn(300, 256, device=cuda_dev) * 4
    p = torch.empty(300, 256, dtype=torch.bfloat16, device=cuda_dev)
    _lib.check(built_lib.its_softmax_rows(p.data_ptr(), s.data_ptr(), 300, 256, _lib.stream_ptr()))
    check_close(p.float(), torch.softmax(s, -1), 5e-3, "softmax")


# ------------------------------------------------------- embeddings ----------
def test_time_embed_and_linear(cuda_dev, built_lib):
    from its_b200 import _lib
    ch, B = 128, 5
    freq = torch.exp(-(torch.arange(0, ch, 2).float() / ch * math.log(10000))).to(cuda_dev)
    t = torch.tensor([0, 1, 999, 1999, 2999], dtype=torch.int64, device=cuda_dev)
    out = torch.empty(B, ch, device=cuda_dev)
    _lib.check(built_lib.its_time_embed(out.data_ptr(), t.data_ptr(), None, freq.data_ptr(), B, ch, _lib.stream_ptr()))
    emb = t.float()[:, None] * freq[None]
    ref = torch.stack([torch.sin(emb), torch.cos(emb)], -1).reshape(B, ch)
    assert (out - ref).abs().max().item() < 2e-4      # fp32 argument up to ~3e3 rad
    t_dev = torch.tensor([999], dtype=torch.int32, device=cuda_dev)
    out1 = torch.empty(1, ch, device=cuda_dev)
    _lib.check(built_lib.its_time_embed(out1.data_ptr(), None, t_dev.data_ptr(), freq.data_ptr(), 1, ch, _lib.stream_ptr()))
    assert torch.equal(out1[0], out[2])
    W = torch.randn(512, ch, device=cuda_dev) / 11
    b = torch.randn(512, device=cuda_dev)
    y = torch.empty(B, 512, device=cuda_dev)
    _lib.check(built_lib.its_linear(y.data_ptr(), out.data_ptr(), W.data_ptr(), b.data_ptr(), B, ch, 512, 1, 1, 0,
                                    _lib.stream_ptr()))
    r = F.linear(F.silu(out), W, b)
    assert (y - F.silu(r)).abs().max().item() < 1e-4
    # row-blocked variant (more than one row): same per-element summation order as the one-row kernel
    for rows_n in (3, 21):
        xr = torch.randn(rows_n, ch, device=cuda_dev)
        yr = torch.empty(rows_n, 512, device=cuda_dev)
        _lib.check(built_lib.its_linear(yr.data_ptr(), xr.data_ptr(), W.data_ptr(), b.data_ptr(), rows_n, ch, 512, 1, 0, 0,
                                        _lib.stream_ptr()))
        y1 = torch.empty(1, 512, device=cuda_dev)
        for i in (0, rows_n - 1):
            _lib.check(built_lib.its_linear(y1.data_ptr(), xr[i:i + 1].contiguous().data_ptr(), W.data_ptr(), b.data_ptr(),
                                            1, ch, 512, 1, 0, 0, _lib.stream_ptr()))
            assert torch.equal(y1[0], yr[i])
        assert (yr - F.linear(F.silu(xr), W, b)).abs().max().item() < 1e-4
    table = torch.randn(11, ch, device=cuda_dev)
    idx = torch.tensor([0, 10, 3, 3, 7], dtype=torch.int64, device=cuda_dev)
    rows = torch.empty(B, ch, device=cuda_dev)
    _lib.check(built_lib.its_embed_rows(rows.data_ptr(), table.data_ptr(), idx.data_ptr(), None, B, ch, 11, _lib.stream_ptr()))
    assert torch.equal(rows, table[idx])


# -------------------------------------------------------- verifiers ----------
def test_verifier_kernels_vs_golden_and_oracle(cuda_dev, built_lib):
    from its_b200.search import verifier as V
    from tests.util import golden
    g = golden("verifier")
    for i, im in enumerate(cases.verifier_images()):
        d = im.to(cuda_dev)
        assert abs(V.OracleVerifier().score(d) - float(g[f"oracle_{i}"])) < 1e-3
        assert abs(V.AestheticPredictor().score(d) - float(g[f"aesthetic_{i}"])) < 1e-3
        a, b = V.SelfSupervisedVerifier().score(d), float(g[f"self_supervised_{i}"])
        assert (math.isnan(a) and math.isnan(b)) or abs(a - b) < 1e-3
        f = V.SelfSupervisedVerifier().extract_features(d)
        assert (f.cpu() - F.adaptive_avg_pool2d(im, (8, 8)).flatten(1)).abs().max().item() < 1e-5
    # per-candidate scores of a population == per-call scores of each candidate
    pop = torch.cat([im for im in cases.verifier_images()[:1]] * 3).to(cuda_dev) * torch.linspace(0.2, 1, 12, device=cuda_dev).view(12, 1, 1, 1)
    for ver, fn in ((V.OracleVerifier(), O.oracle_verifier_score), (V.AestheticPredictor(), O.aesthetic_score),
                    (V.SelfSupervisedVerifier(), O.self_supervised_score)):
        s = ver.score_candidates(pop, 4).cpu()
        for c in range(3):
            assert abs(s[c].item() - fn(pop[4 * c:4 * c + 4].cpu())) < 1e-3
    assert V.OracleVerifier(dataset_stats={"mu": 0}).score(pop) == pytest.approx(pop.mean().item(), abs=1e-5)


def test_argmax_first(cuda_dev, built_lib):
    from its_b200.search.search_algorithm import argmax_first
    nan, inf = float("nan"), float("inf")
    assert argmax_first(torch.tensor([0.5, nan, 0.7, 0.7, 0.1], device=cuda_dev)) == (2, pytest.approx(0.7))
    assert argmax_first(torch.tensor([nan, nan], device=cuda_dev)) == (-1, -inf)
    assert argmax_first(torch.tensor([-inf, -inf], device=cuda_dev)) == (-1, -inf)
    s = torch.randn(70000, device=cuda_dev)
    s[[123, 60000]] = 9.0
    assert argmax_first(s) == (123, 9.0)
    assert argmax_first(torch.tensor([3.0], device=cuda_dev)) == (0, 3.0)


def test_topk_first(cuda_dev, built_lib):
    """its_topk_first: score descending, first index on ties, NaN / -inf never rank, -1 past the eligible ones."""
    from its_b200.search.search_algorithm import argmax_first, topk_first
    v = torch.tensor([0.5, float("nan"), 2.0, 2.0, -1.0, float("-inf"), 0.5, 2.0], device=cuda_dev)
    idx, val = topk_first(v, 8)
    assert idx == [2, 3, 7, 0, 6, 4, -1, -1]
    assert val[:6] == [2.0, 2.0, 2.0, 0.5, 0.5, -1.0]
    g = torch.Generator().manual_seed(3)
    s = torch.randn(5000, generator=g).to(cuda_dev)
    idx, val = topk_first(s, 17)
    ref = torch.sort(s, descending=True, stable=True)
    assert idx == ref.indices[:17].tolist() and val == ref.values[:17].tolist()
    assert idx[0] == argmax_first(s)[0]
