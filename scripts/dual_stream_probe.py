"""Does running the candidate population as K independent sub-batches on K streams (one CUDA graph,
K parallel branches) fill the launch gaps / prologues / epilogues of the single-batch schedule?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from its_b200.Diffusion import UNet
from its_b200.engine import UNetPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = UNet(T=1000, ch=128, ch_mult=[1, 2, 3, 4], attn=[1], num_res_blocks=2, dropout=0.15).to(dev).eval()


def timed(graph, reps=5):
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


import os
for K in [int(k) for k in os.environ.get('KS', '1,2,4').split(',')]:
    plans = [UNetPlan(net, B // K, 32, 32, n_img_in=B // K, uniform_t=True) for _ in range(K)]   # distinct buffers
    for p in plans:
        p.x_in.normal_()
        p.t_dev.fill_(500)
        for _ in range(2):
            p.run()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(K)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for _ in range(2):       # two UNet passes back to back per branch
            for s, p in zip(streams, plans):
                s.wait_stream(main)
                with torch.cuda.stream(s):
                    p.run()
            for s in streams:
                main.wait_stream(s)
    us = timed(g) / 2
    print(f"B={B} as {K} x {B//K}: {us:8.1f} us per UNet pass of the whole population")
