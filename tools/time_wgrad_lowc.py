"""Low-channel weight gradients with and without pixel folding (ops.wgrad_fold).  Usage: python tools/time_wgrad_lowc.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for (H, W, cin, cout) in [(128, 384, 32, 32), (128, 384, 64, 64), (128, 384, 64, 32), (128, 384, 32, 64), (64, 192, 64, 64),
                          (64, 192, 64, 128), (64, 192, 128, 128)]:
    x = torch.randn(B, H, W, cin, device='cuda').half()
    dy = torch.randn(B, H, W, cout, device='cuda').half()
    dw = torch.empty(cout, 9, cin, device='cuda')
    res = []
    for fold in (0, 1):
        ops._WGRAD_FOLD = fold
        for _ in range(2):
            ops.conv_wgrad(x, dy, dw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv_wgrad(x, dy, dw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * H * W * cout * 9 * cin
        res.append(f'fold mode {fold} (f={ops.wgrad_fold(cin, cout, W)}): {ms * 1e3:8.1f} us {fl / ms / 1e9:6.0f} TF/s')
    print(f'wgrad {H}x{W} {cin}->{cout} B={B}: ' + ' | '.join(res))
    del x, dy
