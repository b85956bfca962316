"""Runs the merged transposed conv of the last decoder level a few times (for ncu captures)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, dev = 64, 'cuda'
h, w, cin, cout = 64, 192, 128, 64
x = torch.randn(B, h, w, cin, device=dev).half()
wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
demod = torch.ones(B, cout, device=dev)
raw = torch.zeros(B, 2 * h + 2, 2 * w + 2, cout, device=dev, dtype=torch.float16)
op = ops.convt_s2_merged(x, ops.convt_merged_weight(wt, 1.0), raw, demod)
for _ in range(4):
    op()
torch.cuda.synchronize()
print('ok')
