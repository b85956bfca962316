"""Times the backward pieces of the encoder ConvLayer at B=64 shapes: b200ir_lrelu_bias_bwd (HBM-bound: 6 bytes per
element) and the dgrad conv (b200ir_conv_igemm with adjoint weights).  Usage: python tools/time_backward.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = 64
for (H, W, C) in [(128, 384, 32), (64, 192, 64), (32, 96, 256), (16, 48, 256)]:
    dy = torch.randn(B, H, W, C, device='cuda').half()
    y = torch.randn(B, H, W, C, device='cuda').half()
    dz = torch.empty_like(dy)
    db = torch.empty(C, device='cuda')
    ms = timed(lambda: ops.lrelu_bias_bwd(dy, y, dz, db))
    print(f'lrelu_bias_bwd {H}x{W}x{C}: {ms * 1e3:8.1f} us  {6.0 * dy.numel() / ms / 1e6:7.0f} GB/s')
for (H, W, cin, cout) in [(32, 96, 256, 256), (16, 48, 256, 256), (64, 192, 64, 64), (128, 384, 32, 32)]:
    dz = torch.randn(B, H, W, cout, device='cuda').half()
    wp = (torch.randn(cout, 9 * cin, device='cuda') / (3 * cin ** 0.5)).half()
    dx = torch.empty(B, H, W, cin, device='cuda', dtype=torch.float16)
    op = ops.conv_dgrad(dz, ops.conv_dgrad_weight(wp, cin), dx)
    ms = timed(op)
    print(f'dgrad {H}x{W} {cout}->{cin}: {ms * 1e3:8.1f} us  {2.0 * B * H * W * cout * 9 * cin / ms / 1e9:7.1f} TF/s')
