"""fir_down2_adjoint at the training step's shapes (B = 256): tiled vs per-pixel kernel is selected by the library; prints
the algorithmic bandwidth (d + add in, out).  Usage: python tools/time_fir_adjoint.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for (h, w, c) in [(64, 192, 128), (32, 96, 256), (16, 48, 512), (64, 192, 32), (8, 24, 512), (4, 12, 512)]:
    d = torch.randn(B, h, w, c, device='cuda').half()
    add = torch.randn(B, 2 * h, 2 * w, c, device='cuda').half()
    out = torch.empty_like(add)
    for _ in range(2):
        ops.fir_down2_adjoint(d, out, add=add)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.fir_down2_adjoint(d, out, add=add)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nbytes = (d.numel() + 2 * add.numel()) * 2
    print(f'fir_down2_adjoint B={B} {h}x{w}x{c}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s')
    del d, add, out
