"""Per-CTA timeline of one tap-GEMM launch (clock stamps written by the kernel when
its_conv_desc.dbg is set): where does a CTA's life go — setup, waiting for TMA,
MMA, epilogue?  Usage: python scripts/conv_timeline.py [H] [Cin] [Cout] [B] [cluster]"""
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
cluster = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
w = pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / 30).to(torch.bfloat16).contiguous()
plan = UNetPlan.scratch(dev, B, 0)
plan.split_k = False
out = plan.conv([(x, Cin, 0, 1, False)], [(taps_square(3), 0, 0, 0)], H, H, w, Cout)
d = plan.descs[0]
d.cluster = cluster
n_cta = 8192
dbg = torch.zeros(n_cta, 64, dtype=torch.int64, device=dev)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
d.dbg = dbg.data_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(); e1.record()
torch.cuda.synchronize()
t = dbg.cpu().numpy()
used = t[:, 0] != 0
t = t[used]
print(f"H={H} Cin={Cin} Cout={Cout} B={B} cluster={cluster}: {used.sum()} CTAs, kernel {e0.elapsed_time(e1)*1e3:.1f} us")
g0 = t[:, 0].min()
nkb = int((t[0, 8:] != 0).sum())
clk = 1.9  # GHz (approx, for display)
print("cta  sm  start_us  setup_us  first_full_us  mainloop_us  per_kb_ns  epilogue_us  total_us")
order = np.argsort(t[:, 0])
for i in list(order[:6]) + list(order[len(order)//2:len(order)//2+4]) + list(order[-4:]):
    r = t[i]
    c0 = r[1]
    setup = (r[2] - c0) / clk / 1e3
    first = (r[8] - c0) / clk / 1e3
    last = (r[8 + nkb - 1] - c0) / clk / 1e3
    acc = (r[4] - c0) / clk / 1e3
    end = (r[5] - c0) / clk / 1e3
    print(f"{i:4d} {r[6]:3d} {(r[0]-g0)/1e3:9.2f} {setup:8.2f} {first:13.2f} {acc-first:11.2f} {(last-first)/(max(nkb-1,1))*1e3:9.0f} {end-acc:11.2f} {end:9.2f}")
kb = (t[:, 8:8 + nkb] - t[:, 1:2]) / clk          # ns since CTA entry
print("median ns between consecutive full barriers:", np.median(np.diff(kb, axis=1), axis=0).round(0))
print("median: setup %.2f us, first full %.2f us, acc complete %.2f us, end %.2f us" % tuple(
    np.median((t[:, j] - t[:, 1]) / clk / 1e3) for j in (2, 8, 4, 5)))
print("CTA start spread (us): p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile((t[:, 0] - g0) / 1e3, [50, 90, 100])))
