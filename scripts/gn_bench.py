"""Time the GroupNorm launches (apply-from-statistics and the self-contained cluster kernel) on the
UNet's tensor shapes: CUDA events around a graph of 20 launches.  Usage: python scripts/gn_bench.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200.engine import UNetPlan, pack_conv_weight, taps_square

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
for H, C0, C1 in [(32, 128, 0), (32, 128, 128), (32, 256, 128), (16, 256, 0), (16, 256, 256), (8, 384, 384), (4, 512, 512)]:
    for mode in ("stats", "cluster"):
        plan = UNetPlan.scratch(dev, B, 0)
        plan.split_k, plan.schedule = False, 2
        outs, keep = [], []
        for C in [c for c in (C0, C1) if c]:
            x = torch.randn(B, H, H, 64, device=dev).to(torch.bfloat16)
            w = pack_conv_weight(torch.randn(C, 64, 3, 3, device=dev) / 24).to(torch.bfloat16).contiguous()
            keep += [x, w]
            outs.append(plan.conv([(x, 64, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, w, C, want_stats=(mode == "stats")))
        gn = torch.nn.GroupNorm(32, C0 + C1).to(dev)
        plan.run()
        n0 = len(plan.ops)
        plan.group_norm(outs, gn, True)
        gn_ops = plan.ops[n0:]
        plan.ops = gn_ops
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                plan.run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        mb = B * H * H * (C0 + C1) * 4 / 1e6
        print(f"H={H:2d} C={C0}+{C1} {mode:8s} {plan.op_info[-1][0]:18s}: {us:6.1f} us  {mb/us*1e-3:6.2f} TB/s (read+write {mb:.1f} MB)")
