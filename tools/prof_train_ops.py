"""Launches the memory-bound training kernels at their largest shapes of the B = 64 plate step (128x384, 64 channels) so that
`ncu --set full -k regex:"sft_mod_bwd|style_act_bwd|to_rgb_bwd"` can capture them.  Usage: python tools/prof_train_ops.py"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from image_restoration_b200 import ops  # noqa: E402

B, h, w, C, c_sft = 64, 128, 384, 64, 32
dev = 'cuda'
torch.manual_seed(0)
a = torch.randn(B, h, w, C, device=dev).half()
g = torch.randn(B, h, w, C, device=dev).half()
sc = (1 + 0.1 * torch.randn(B, h, w, c_sft, device=dev)).half()
sh = torch.randn(B, h, w, c_sft, device=dev).half()
s = 1 + 0.1 * torch.randn(B, C, device=dev)
da, dsc, dsh = torch.empty_like(a), torch.empty_like(sc), torch.empty_like(sh)
ds = torch.zeros(B, C, device=dev)
nz = torch.randn(B, 1, h, w, device=dev)
gain, bias = torch.tensor([0.2], device=dev), torch.randn(C, device=dev)
dd = torch.zeros(B, C, device=dev)
drgb = torch.randn(B, 3, h, w, device=dev)
wr = torch.randn(3, C, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, nbytes in (
        ('sft_mod_bwd', lambda: ops.sft_mod_bwd(g, a, sc, sh, s, da, False, dsc, dsh, ds), (3 * a.numel() + 4 * sc.numel()) * 2),
        ('style_act_bwd', lambda: ops.style_act_bwd(g, a, nz, gain, bias, s, 4.0, da, dd), 3 * a.numel() * 2 + nz.numel() * 4),
        ('to_rgb_bwd', lambda: ops.to_rgb_bwd(drgb, a, wr, s, da, False, ds), 2 * a.numel() * 2 + drgb.numel() * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(10):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f'{name}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s algorithmic')
