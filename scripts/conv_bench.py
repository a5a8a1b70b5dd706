"""Time single tap-GEMM launches (CUDA events, 20 reps after 3 warm-ups) for the UNet's layer
shapes under a given schedule / N-tile override.
Usage: python scripts/conv_bench.py [schedule=0|1|2] [bn=0|64|128|192|256] [B=64]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200.engine import UNetPlan, pack_conv_weight, taps_square

schedule = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
cluster = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # persistent schedule: 1 = one row box per tile
dev = torch.device("cuda:0")
SHAPES = [  # H, Cin, Cout, k
    (32, 128, 128, 3), (32, 256, 128, 3), (32, 384, 128, 3),
    (16, 256, 256, 3), (16, 512, 256, 3), (16, 256, 256, 1), (16, 256, 512, 1),
    (8, 384, 384, 3), (8, 768, 384, 3), (8, 256, 256, 3),
    (4, 512, 512, 3), (4, 1024, 512, 3),
]
print(f"schedule={schedule} bn={bn} B={B} cluster={cluster}")
for H, Cin, Cout, k in SHAPES:
    x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
    w = pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / 30).to(torch.bfloat16).contiguous()
    bias = torch.randn(Cout, device=dev)
    plan = UNetPlan.scratch(dev, B, 0)
    plan.schedule = schedule
    plan.conv([(x, Cin, 0, 1, False)], [(taps_square(k), 0, 0, 0)], H, H, w, Cout, bias=bias)
    d = plan.descs[0]
    if bn and Cout % bn == 0 and d.splits <= 1:
        d.bn = bn
    d.cluster = cluster
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            plan.run()
    g.replay()
    torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    flops = 2 * B * H * H * Cout * Cin * k * k
    print(f"H={H:2d} Cin={Cin:4d} Cout={Cout:3d} k={k} bn={d.bn} splits={d.splits} stats={'y' if d.stats else 'n'}: {us:7.1f} us  {flops/us/1e6:7.1f} TFLOP/s")
