"""Runs b200ir_first_conv and the largest b200ir_rgb_combine of the B = 64 step a few times (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, H, W, C = 64, 128, 384, 32
x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
w = torch.randn(C, 3, device='cuda')
b = torch.zeros(C, device='cuda')
out = torch.empty(B, H, W, C, device='cuda', dtype=torch.float16)
for _ in range(3):
    ops.first_conv(x, w, b, out)
torch.cuda.synchronize()
print('ok')
