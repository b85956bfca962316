"""Forward + backward time of the five encoder ResBlocks of the 128x384 network at batch B (default 64) through
image_restoration_b200.backward.ResBlockFunction (CUDA events around fwd and around bwd, 5 repetitions after warm-up).
Algorithmic FLOPs: forward 2 * MACs of conv1 + conv2 + skip; backward twice that (dgrad + wgrad).
Usage: python tools/time_resblock.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200.backward import res_block  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tot_f = tot_b = 0.0
for (H, W, cin, cout) in [(128, 384, 32, 64), (64, 192, 64, 256), (32, 96, 256, 256), (16, 48, 256, 256), (8, 24, 256, 256)]:
    torch.manual_seed(0)
    par = [torch.randn(cin, cin, 3, 3), torch.zeros(cin), torch.randn(cout, cin, 3, 3), torch.zeros(cout), torch.randn(cout, cin, 1, 1)]
    par = [p.cuda().requires_grad_() for p in par]
    x = torch.randn(B, H, W, cin, device='cuda').half().requires_grad_()
    dout = torch.randn(B, H // 2, W // 2, cout, device='cuda').half()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(7):
        ev[0].record()
        out = res_block(x, *par)
        ev[1].record()
        out.backward(dout)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf += ev[0].elapsed_time(ev[1]) / 5
            tb += ev[1].elapsed_time(ev[2]) / 5
    fl = 2.0 * B * (H * W * 9 * cin * cin + (H // 2) * (W // 2) * (9 * cin * cout + cin * cout))
    print(f'ResBlock {H}x{W} {cin}->{cout}: fwd {tf * 1e3:7.0f} us ({fl / tf / 1e9:6.0f} TF/s)  bwd {tb * 1e3:7.0f} us '
          f'({2 * fl / tb / 1e9:6.0f} TF/s)')
    tot_f += tf
    tot_b += tb
print(f'encoder ResBlocks B={B}: fwd {tot_f:.2f} ms, bwd {tot_b:.2f} ms')
