"""UNet-pass throughput of the other BASELINE.json configurations (C: conditional CFG net, E: 64x64
net), one pass of the launch plan in a CUDA graph; candidate images/s extrapolated over the T steps
(x2 UNet evaluations per step for classifier-free guidance, batched as one 2B pass)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--candidates", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
peak = 1403.3e12
out = []
for name in ("A", "C", "E"):
    if name == "C":
        from its_b200.DiffusionFreeGuidence import UNet
        net = UNet(T=1000, num_labels=10, ch=128, ch_mult=[1, 2, 3, 4], num_res_blocks=2, dropout=0.15).to(dev).eval()
        img, T, n_img, n_in = 32, 1000, 2 * a.candidates, a.candidates
    else:
        from its_b200.Diffusion import UNet
        img, T = (32, 1000) if name == "A" else (64, 2000)
        net = UNet(T=T, ch=128, ch_mult=[1, 2, 3, 4], attn=[1] if img == 32 else [2], num_res_blocks=2, dropout=0.15).to(dev).eval()
        n_img, n_in = a.candidates, a.candidates
    plan = net.plan(n_img, img, img, n_img_in=n_in, uniform_t=True)
    plan.x_in.normal_()
    plan.t_dev.fill_(T // 2)
    if plan.labels is not None:
        plan.labels.copy_(torch.cat([1 + torch.arange(a.candidates) % 10, torch.zeros(a.candidates, dtype=torch.long)]).to(dev))
    plan.run_label_ops()     # label embedding -> cond_proj: once per trajectory (labels never change inside one)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(3):
            plan.run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rec = {"config": name, "img": img, "T": T, "candidates": a.candidates, "unet_images_per_pass": n_img,
           "launches_per_pass": plan.n_launches, "ms_per_unet_pass": ms, "gflop_per_pass": plan.flops / 1e9,
           "tflops": plan.flops / ms / 1e9, "frac_of_sustained_bf16_peak": plan.flops / (ms * 1e-3) / peak,
           "candidate_images_per_s_extrapolated": a.candidates / (ms * 1e-3 * T)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del plan, net
    torch.cuda.empty_cache()
