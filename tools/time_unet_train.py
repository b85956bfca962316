"""Forward + backward time of the trainable part of GFPGANv1OCR (U-Net encoder -> style code, decoder -> SFT conditions)
through image_restoration_b200.backward.unet_forward at batch B (default 64), stock-init weights, CUDA events.
Usage: python tools/time_unet_train.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.backward import unet_forward  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW)
names = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale', 'condition_shift')
sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in names}
n_par = sum(v.numel() for v in sd.values())
x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
cots = None
for it in range(6):
    for v in sd.values():
        v.grad = None
    ev[0].record()
    style, conds = unet_forward(sd, x, different_w=True, num_style_feat=NET_KW['num_style_feat'])
    ev[1].record()
    if cots is None:
        cots = [torch.randn_like(t) for t in [style] + conds]
    torch.autograd.backward([style] + conds, cots)
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 2:
        tf += ev[0].elapsed_time(ev[1]) / 4
        tb += ev[1].elapsed_time(ev[2]) / 4
print(f'U-Net (trainable part, {n_par / 1e6:.1f} M parameters) B={B}: forward {tf:.2f} ms, backward {tb:.2f} ms, '
      f'{B / (tf + tb) * 1e3:.0f} crops/s fwd+bwd; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
