import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import ddpm_oracle as O
from tests.util import build_shell, rel_err, rms_err
from tests import cases
dev = torch.device("cuda:0")
for kind, ch, mult, img, B in [("uncond", 32, [1, 2], 8, 3), ("cond", 32, [1, 2], 8, 3), ("uncond", 96, [1, 2, 2], 16, 2),
                               ("cond", 160, [1, 2], 16, 2), ("uncond", 128, [1, 1, 2, 2], 32, 2)]:
    cfg = dict(kind=kind, T=50, ch=ch, ch_mult=mult, attn=[1], num_res_blocks=1, dropout=0.0, num_labels=10,
               weight_seed=51, img=img, B=B, input_seed=151)
    try:
        net, sd = build_shell(cfg, dev)
        x, t, labels = cases.forward_inputs(cfg)
        args = (x.to(dev), t.to(dev)) + ((labels.to(dev),) if labels is not None else ())
        eps = net(*args).cpu()
        with torch.no_grad():
            ref = O.unet_forward(sd, x, t, labels)
        plan = next(iter(net._plans.values()))
        kinds = sorted(set(k for k, _, _ in plan.op_info))
        print(kind, ch, mult, img, "rms %.2e rel %.2e" % (rms_err(eps, ref), rel_err(eps, ref)), kinds)
    except Exception as e:
        print(kind, ch, mult, img, "FAILED", type(e).__name__, str(e)[:300])
