"""Per-CTA timeline of one persistent tap-GEMM launch (its_conv_desc.dbg stamps of
conv_persist_sm100.cu): setup, per-tile accumulator-complete and epilogue-done times.
Usage: python scripts/persist_timeline.py [H] [Cin] [Cout] [B] [bn]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from its_b200.engine import UNetPlan, pack_conv_weight, taps_square

H = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Cin = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cout = int(sys.argv[3]) if len(sys.argv) > 3 else 128
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
split = int(sys.argv[6]) if len(sys.argv) > 6 else 0
force = int(sys.argv[7]) if len(sys.argv) > 7 else 0       # forced split-K factor (with bn)
dev = torch.device("cuda:0")
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
w = pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / 30).to(torch.bfloat16).contiguous()
bias = torch.randn(Cout, device=dev)
plan = UNetPlan.scratch(dev, B, 0)
plan.split_k, plan.schedule = bool(split), 2
if force:
    plan._persist_plan = lambda *a, **k: (bn, force)
out = plan.conv([(x, Cin, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, w, Cout, bias=bias)
d = plan.descs[0]
if bn and not force:
    d.bn = bn
print('bn', d.bn, 'splits', d.splits)
dbg = torch.zeros(256, 64, dtype=torch.int64, device=dev)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
d.dbg = dbg.data_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(); e1.record()
torch.cuda.synchronize()
t = dbg.cpu().numpy()
t = t[t[:, 0] != 0]
clk = 1.9
print(f"H={H} Cin={Cin} Cout={Cout} B={B} bn={bn}: {len(t)} CTAs, kernel {e0.elapsed_time(e1)*1e3:.1f} us (eager, incl. launch)")
g = torch.cuda.CUDAGraph()
d.dbg = None
with torch.cuda.graph(g):
    for _ in range(20):
        plan.run()
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"steady state (20 launches in a graph): {e0.elapsed_time(e1)*1e3/20:.2f} us per launch")
g0 = t[:, 0].min()
ntile = int((t[0, 8:16] != 0).sum())
print("cta sm start_us setup_us | per tile: acc_complete_us / epilogue_done_us ... | end_us")
order = np.argsort(t[:, 0])
for i in list(order[:4]) + list(order[len(order)//2:len(order)//2+2]) + list(order[-4:]):
    r = t[i]
    c0 = r[1]
    us = lambda v: (v - c0) / clk / 1e3
    tiles = " ".join(f"{us(r[8+j]):6.2f}/{us(r[24+j]):6.2f}" for j in range(8) if r[8 + j])
    print(f"{i:3d} {r[6]:3d} {(r[0]-g0)/1e3:6.2f} {us(r[2]):5.2f} | {tiles} | {us(r[3]):6.2f}")
med = lambda col: np.median((t[:, col] - t[:, 1]) / clk / 1e3)
print("median: setup %.2f, first acc complete %.2f, first epilogue done %.2f, end %.2f us" % (med(2), med(8), med(24), med(3)))
print("CTA start spread (us): p50 %.2f max %.2f" % tuple(np.percentile((t[:, 0] - g0) / 1e3, [50, 100])))
