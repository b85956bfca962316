"""Times b200ir_conv_wgrad at the 3x3 layers of the B=64 forward that it supports (cin % 64 == 0, cout % 128 == 0).
Usage: python tools/time_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B = 64
for (H, W, cin, cout) in [(32, 96, 256, 256), (32, 96, 256, 512), (32, 96, 512, 512), (16, 48, 256, 256), (16, 48, 512, 512),
                          (64, 192, 128, 128), (64, 192, 64, 128), (8, 24, 256, 256), (4, 12, 256, 256)]:
    x = torch.randn(B, H, W, cin, device='cuda').half()
    dy = torch.randn(B, H, W, cout, device='cuda').half()
    dw = torch.empty(cout, 9, cin, device='cuda')
    for _ in range(2):
        ops.conv_wgrad(x, dy, dw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv_wgrad(x, dy, dw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * W * cout * 9 * cin
    print(f'wgrad {H}x{W} {cin}->{cout}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TF/s')
