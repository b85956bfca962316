"""Latency / throughput of the forward at small batches (api.py serves one crop per request).
Usage: python tools/time_batches.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402

torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
for B in (1, 2, 4, 8, 16, 32, 64, 128):
    x = (torch.rand(B, 3, H, W) * 2 - 1).cuda()
    with torch.no_grad():
        for _ in range(3):
            net(x, return_rgb=False, randomize_noise=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            net(x, return_rgb=False, randomize_noise=False)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f'B={B:4d}: {ms:7.3f} ms / forward   {B / ms * 1e3:8.0f} crops/s')
