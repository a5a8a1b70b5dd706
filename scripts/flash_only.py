"""Micro-benchmark of the streaming attention kernel alone (128 images, N = 1024 tokens, C = 128): 10 back-to-back
launches timed with CUDA events; also the command profiled in profiles/r01_ncu_full_attention_flash.txt."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from its_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
B, N, C = 128, 1024, 128
qkv = (torch.randn(B, N, 3 * C, device=dev)).to(torch.bfloat16)
out = torch.empty(B, N, C, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(C, device=dev)
for _ in range(3):
    _lib.check(L.its_attention_flash(out.data_ptr(), qkv.data_ptr(), None, bias.data_ptr(), B, N, C, C ** -0.5, _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.its_attention_flash(out.data_ptr(), qkv.data_ptr(), None, bias.data_ptr(), B, N, C, C ** -0.5, _lib.stream_ptr())
e1.record(); torch.cuda.synchronize()
print("flash us", e0.elapsed_time(e1) * 100)
