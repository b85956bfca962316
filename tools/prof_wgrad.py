"""Runs b200ir_conv_wgrad (and the lrelu / bias backward) a few times at one B = 64 layer shape, for ncu captures.
Usage: python tools/prof_wgrad.py H W cin cout"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

H, W, cin, cout = map(int, sys.argv[1:5])
x = torch.randn(64, H, W, cin, device='cuda').half()
dy = torch.randn(64, H, W, cout, device='cuda').half()
y = torch.randn(64, H, W, cout, device='cuda').half()
dw = torch.empty(cout, 9, cin, device='cuda')
for _ in range(3):
    ops.lrelu_bias_bwd(dy, y)
    ops.conv_wgrad(x, dy, dw)
torch.cuda.synchronize()
print('ok')
