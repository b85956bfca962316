"""Sweeps the split-K factor of b200ir_conv_wgrad (B200IR_WGRAD_SPLITS) per layer shape.  Usage: python tools/sweep_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B = 64
for (H, W, cin, cout) in [(32, 96, 256, 256), (32, 96, 512, 512), (16, 48, 256, 256), (64, 192, 128, 128), (64, 192, 64, 128),
                          (8, 24, 256, 256), (4, 12, 256, 256)]:
    x = torch.randn(B, H, W, cin, device='cuda').half()
    dy = torch.randn(B, H, W, cout, device='cuda').half()
    dw = torch.empty(cout, 9, cin, device='cuda')
    nc = 2 if cin % 128 == 0 else 1
    units = (cout // 128) * (cin // (64 * nc)) * 3
    tiles = B * -(-W // 32) * -(-H // 4)
    line = f'{H}x{W} {cin}->{cout} units={units} tiles={tiles}:'
    for ctas in (37, 74, 111, 148, 222, 296, 444, 592):
        s = max(1, ctas // units)
        os.environ['B200IR_WGRAD_SPLITS'] = str(s)
        for _ in range(2):
            ops.conv_wgrad(x, dy, dw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv_wgrad(x, dy, dw)
        e1.record()
        torch.cuda.synchronize()
        line += f'  s{s}({s * units}):{e0.elapsed_time(e1) * 100:.0f}us'
    print(line)
