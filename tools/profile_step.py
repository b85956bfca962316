"""Runs a few eager (non-graph) forward steps at batch B so that `ncu` sees every kernel launch of one step.
Usage: python tools/profile_step.py [B] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
eng = net.engine()
eng.use_graphs = False
x = (torch.rand(B, 3, H, W) * 2 - 1).cuda()
for _ in range(steps):
    y, _ = net(x, return_rgb=False, randomize_noise=False)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
