"""Experiment: final_linear (12288 -> 3072, 64 rows) as a 1x1 conv with different N-tile widths (CTA counts).
Usage: python tools/time_linear.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, K, N = 64, 12288, 3072
x = torch.randn(B, K, device='cuda').half()
w = (torch.randn(N, K, device='cuda') / math.sqrt(K)).half()
bias = torch.zeros(N, device='cuda')
ref = x.float() @ w.float().t()
for bn in (256, 128, 64, 32, 16):
    out = torch.empty(B, N, device='cuda')
    op = ops.linear_as_conv(x, w, out, bias=bias, block_n=bn)
    op()
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f'block_n={bn:3d}: {us:6.1f} us  {N * K * 2 / us / 1e6:6.2f} TB/s weight stream  max err {err:.3e}')
