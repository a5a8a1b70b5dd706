"""k-block cost of the persistent kernel under forced (N tile, split-K) choices on the small maps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200.engine import UNetPlan, pack_conv_weight, taps_square

dev = torch.device("cuda:0")
B = 64
ws = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
CASES = [  # H, Cin, Cout, bn, splits
    (4, 512, 512, 128, 1), (4, 512, 512, 64, 1), (4, 512, 512, 64, 2), (4, 1024, 512, 64, 1), (4, 1024, 512, 64, 2),
    (8, 384, 384, 192, 1), (8, 384, 384, 128, 1), (8, 384, 384, 64, 1), (8, 768, 384, 128, 1),
    (8, 256, 256, 256, 1), (8, 256, 256, 128, 1), (8, 256, 256, 64, 1), (16, 128, 128, 128, 1), (16, 128, 128, 64, 1),
]
for H, Cin, Cout, bn, S in CASES:
    x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
    w = pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / 30).to(torch.bfloat16).contiguous()
    plan = UNetPlan.scratch(dev, B, 0)
    plan.split_k, plan.schedule = False, 2
    plan.conv([(x, Cin, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, w, Cout, want_stats=False)
    d = plan.descs[0]
    d.bn, d.cluster = bn, 1
    if S > 1:
        d.splits, d.ws, d.ws_elems = S, ws.data_ptr(), ws.numel()
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
    nkb = 9 * Cin // 64
    tiles = plan._tiles_at(B, H, H) * (Cout // bn)
    print(f"H={H} Cin={Cin} Cout={Cout} bn={bn} S={S} items={tiles*S:3d} kb/item={nkb//S:3d}: {us:6.1f} us -> {(us-7.0)/(nkb/S)*1e3:5.0f} ns/kb")
