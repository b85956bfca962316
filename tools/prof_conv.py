"""Runs one conv layer config of tools/time_conv.py a few times (for ncu captures).
Usage: python tools/prof_conv.py H W cin cout kind"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import time_conv  # noqa: E402

H, W, cin, cout = map(int, sys.argv[1:5])
op, keep = time_conv.build(H, W, cin, cout, sys.argv[5])
for _ in range(4):
    op()
torch.cuda.synchronize()
print('ok')
