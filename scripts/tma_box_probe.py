"""Does the A-operand box shape (resolution) or the zero-padding (3x3 vs 1x1 taps) change the cost of a
k-block?  Same number of tiles (64), same N tile (256), different feature-map sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200.engine import UNetPlan, pack_conv_weight, taps_square

dev = torch.device("cuda:0")
ntile = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for k, Cin in ((3, 512), (1, 1024)):
    for H in (32, 16, 8, 4):
        B = ntile * 128 // (H * H)
        Cout = 256
        x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
        w = pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / 30).to(torch.bfloat16).contiguous()
        plan = UNetPlan.scratch(dev, B, 0)
        plan.split_k, plan.schedule = False, 2
        plan.conv([(x, Cin, 0, 1, False)], [(taps_square(k), 0, 0, 0)], H, H, w, Cout, want_stats=False)
        d = plan.descs[0]
        d.bn, d.cluster = 256, 1
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
        nkb = k * k * Cin // 64
        print(f"k={k} H={H:2d} B={B:4d} tiles={ntile} nkb={nkb}: {us:6.1f} us  -> {(us-7.0)/nkb*1e3:5.0f} ns per k-block (7 us fixed assumed)")
