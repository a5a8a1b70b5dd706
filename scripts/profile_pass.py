"""One UNet pass (config A, 64 candidates) for ncu: `--passes P` eager passes of the
launch plan, printing the launch list (kind, algorithmic FLOPs) so that ncu's
per-launch durations can be joined with it.

    python scripts/profile_pass.py --passes 3 --list > gpurun_out/launch_kinds.txt
    ncu --metrics gpu__time_duration.sum --clock-control none -s <2 passes> -c <1 pass> ... python scripts/profile_pass.py --passes 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from its_b200.Diffusion import UNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--img", type=int, default=32)
ap.add_argument("--list", action="store_true")
ap.add_argument("--time", action="store_true", help="CUDA-event time per op (no profiler)")
a = ap.parse_args()

torch.manual_seed(0)
dev = torch.device("cuda:0")
net = UNet(T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1] if a.img == 32 else [2], num_res_blocks=2, dropout=0.15).to(dev).eval()
plan = net.plan(a.batch, a.img, a.img, n_img_in=a.batch, uniform_t=True)
plan.x_in.normal_()
plan.t_dev.fill_(500)
if a.list:
    print("launches_per_pass", plan.n_launches)
    for i, ((fn, args), (kind, flops, launches)) in enumerate(zip(plan.ops, plan.op_info)):
        extra = ""
        if kind.startswith("tapgemm"):
            d = args[0]._obj
            extra = (f"B={d.B} Hm={d.Hm} Wm={d.Wm} Cout={d.Cout} K={sum(d.src[d.phase[0].src[t]].C for t in range(d.phase[0].ntaps))}"
                     f" phases={d.nphases} stride={d.src[0].stride}")
        print(i, kind, launches, flops, extra)
for _ in range(a.passes):
    plan.run()
torch.cuda.synchronize()
if a.time:
    from its_b200 import _lib
    s = _lib.stream_ptr()
    evs = []
    for (fn, args), info in zip(plan.ops, plan.op_info):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*args, s); e1.record()
        evs.append((e0, e1, info))
    torch.cuda.synchronize()
    tot = {}
    for i, (e0, e1, (kind, flops, launches)) in enumerate(evs):
        ms = e0.elapsed_time(e1)
        tot.setdefault(kind, [0.0, 0, 0])
        tot[kind][0] += ms; tot[kind][1] += flops; tot[kind][2] += launches
        tf = flops / ms / 1e9 if ms > 0 else 0
        print(f"op {i:3d} {kind:18s} {ms*1e3:9.1f} us  {tf:8.1f} TFLOP/s")
    for k, (ms, fl, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"TOTAL {k:18s} {ms:8.3f} ms  launches {n:4d}  {fl/ms/1e9 if ms else 0:8.1f} TFLOP/s")
print("done")
