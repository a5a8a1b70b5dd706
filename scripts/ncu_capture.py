"""Launches of one UNet pass for Nsight Compute.

    python scripts/ncu_capture.py --config A --list            # op index, kind, launches (no profiler range)
    ncu --profile-from-start off ... python scripts/ncu_capture.py --config A             # one whole pass in the range
    ncu --profile-from-start off ... python scripts/ncu_capture.py --config A --ops 5,7   # only these ops in the range

Configs: A = Model.UNet on 32x32, 64 images; C = ModelCondition.UNet on 32x32, 128 images (64 guided candidates);
E = Model.UNet attn=[2] on 64x64, 64 images.  Three warm passes run first (outside the profiler range), so the
profiled launches see real inputs; ncu itself serialises them and flushes caches, so compare shares, not absolutes.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from its_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="A", choices=["A", "C", "E"])
ap.add_argument("--list", action="store_true")
ap.add_argument("--ops", default="")
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda:0")
if a.config == "C":
    from its_b200.DiffusionFreeGuidence import UNet
    net = UNet(T=1000, num_labels=10, ch=128, ch_mult=[1, 2, 3, 4], num_res_blocks=2, dropout=0.15).to(dev).eval()
    batch, img = 128, 32
else:
    from its_b200.Diffusion import UNet
    img = 32 if a.config == "A" else 64
    net = UNet(T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1] if img == 32 else [2], num_res_blocks=2, dropout=0.15).to(dev).eval()
    batch = 64
# synthetic O(1) weights: the zero-gain initialisers would feed the kernels denormal-sized activations
with torch.no_grad():
    g = torch.Generator().manual_seed(1)
    for p in net.parameters():
        if p.dim() >= 2 and float(p.abs().max()) < 1e-3:
            fan = p[0].numel()
            p.copy_((torch.randn(p.shape, generator=g) / fan ** 0.5).to(dev))
net.invalidate_plans()
plan = net.plan(batch, img, img, n_img_in=batch, uniform_t=True)
plan.x_in.normal_()
plan.t_dev.fill_(500)
if plan.labels is not None:
    plan.labels.copy_((torch.arange(batch) % 11).to(dev))
plan.run_label_ops()
if a.list:
    for i, ((fn, args), (kind, flops, launches)) in enumerate(zip(plan.ops, plan.op_info)):
        extra = ""
        if kind.startswith("tapgemm"):
            d = args[0]._obj
            K = sum(d.src[d.phase[0].src[t]].C for t in range(d.phase[0].ntaps))
            extra = f"H={d.Hm} Cout={d.Cout} K={K} ph={d.nphases} bn={d.bn} S={d.splits} fusedGN={int(bool(d.gn_out))}"
        print(i, kind, launches, flops, extra)
    sys.exit(0)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
s = _lib.stream_ptr()
ops = [int(x) for x in a.ops.split(",") if x] if a.ops else list(range(len(plan.ops)))
torch.cuda.profiler.start()
for i in ops:
    fn, args = plan.ops[i]
    _lib.check(fn(*args, s), fn.__name__)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled ops", len(ops))
