"""Do the small-map levels (8x8 and 4x4: everything from the DownSample into 8x8 to the UpSample out of it)
run faster as K independent sub-populations on K streams?  Times only that section of the launch plan:
one plan of B images against K plans of B/K images as parallel branches of one CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200 import _lib
from its_b200.Diffusion import UNet
from its_b200.engine import UNetPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = UNet(T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1], num_res_blocks=2, dropout=0.15).to(dev).eval()


def section(plan):
    """indices [lo, hi) of the ops whose GEMM rows are 8x8 or 4x4 maps (plus the launches between them)"""
    idx = [i for i, ((fn, a), (kind, _, _)) in enumerate(zip(plan.ops, plan.op_info))
           if kind.startswith("tapgemm") and a[0]._obj.Hm <= 8 and a[0]._obj.Hm > 1]
    return min(idx), max(idx) + 1


def timed(graph, reps=10):
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for K in (1, 2, 4):
    plans = [UNetPlan(net, B // K, 32, 32, n_img_in=B // K, uniform_t=True) for _ in range(K)]
    for p in plans:
        p.x_in.normal_()
        p.t_dev.fill_(500)
        for _ in range(2):
            p.run()
    torch.cuda.synchronize()
    lo, hi = section(plans[0])
    streams = [torch.cuda.Stream() for _ in range(K)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for s, p in zip(streams, plans):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                sp = s.cuda_stream
                for fn, a in p.ops[lo:hi]:
                    fn(*a, sp)
        for s in streams:
            main.wait_stream(s)
    print(f"ops [{lo},{hi}) of {len(plans[0].ops)}: B={B} as {K} x {B // K}: {timed(g):8.1f} us for the section")
