"""Steady-state time of every launch of one UNet pass: each op is replayed 20x back to back inside
its own CUDA graph (same dependency/launch-gap regime as the sampler's step graph), CUDA events
around the replay.  Prints per-op times, per-kind totals and the whole-pass graph time."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200 import _lib
from its_b200.Diffusion import UNet

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--img", type=int, default=32)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--cond", action="store_true", help="config C: ModelCondition.UNet, batch = 2 x candidates (CFG)")
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda:0")
if a.cond:
    from its_b200.DiffusionFreeGuidence import UNet as CUNet
    net = CUNet(T=1000, num_labels=10, ch=128, ch_mult=[1, 2, 3, 4], num_res_blocks=2, dropout=0.15).to(dev).eval()
else:
    net = UNet(T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1] if a.img == 32 else [2], num_res_blocks=2, dropout=0.15).to(dev).eval()
plan = net.plan(a.batch, a.img, a.img, n_img_in=a.batch, uniform_t=True)
plan.x_in.normal_()
plan.t_dev.fill_(500)
if plan.labels is not None:
    plan.labels.copy_((torch.arange(a.batch) % 11).to(dev))
plan.run_label_ops()
for _ in range(3):
    plan.run()
torch.cuda.synchronize()


def timed(fn, reps):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


s = _lib.stream_ptr
tot = collections.defaultdict(lambda: [0.0, 0, 0])
for i, ((fn, args), (kind, flops, launches)) in enumerate(zip(plan.ops, plan.op_info)):
    us = timed(lambda: fn(*args, s()), a.reps)
    extra = ""
    if kind.startswith("tapgemm"):
        d = args[0]._obj
        K = sum(d.src[d.phase[0].src[t]].C for t in range(d.phase[0].ntaps))
        extra = f"H={d.Hm} Cout={d.Cout} K={K} ph={d.nphases} bn={d.bn} S={d.splits} sched={d.schedule}"
    print(f"op {i:3d} {kind:18s} {us:7.1f} us  {flops/us/1e6 if flops else 0:7.1f} TF/s  {extra}")
    tot[kind][0] += us; tot[kind][1] += flops; tot[kind][2] += launches
for k, (us, fl, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"TOTAL {k:18s} {us:8.1f} us  launches {n:4d}  {fl/us/1e6 if fl else 0:7.1f} TF/s")
print(f"SUM of ops {sum(v[0] for v in tot.values()):.1f} us; whole pass in one graph: {timed(plan.run, 5):.1f} us; flops/pass {plan.flops/1e9:.1f} G")
