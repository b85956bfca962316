"""One upfir_act + one fir_pad22 launch at the largest B=64 shapes (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, dev = 64, 'cuda'
h2, w2, C = 128, 384, 64
raw = torch.randn(B, h2 + 2, w2 + 2, C, device=dev).half()
out = torch.empty(B, h2, w2, C, device=dev, dtype=torch.float16)
noise = torch.randn(B, 1, h2, w2, device=dev)
gain, bias = torch.zeros(1, device=dev), torch.zeros(C, device=dev)
sc = torch.randn(B, h2, w2, C // 2, device=dev).half()
sh = torch.randn(B, h2, w2, C // 2, device=dev).half()
sn = torch.ones(B, C, device=dev)
x = torch.randn(B, 128, 384, 32, device=dev).half()
p = torch.zeros(B, 130, 386, 32, device=dev, dtype=torch.float16)
for _ in range(3):
    ops.upfir_act(raw, out, noise, h2 * w2, gain, bias, sc, sh, C // 2, sn)
    ops.fir_pad22(x, p)
torch.cuda.synchronize()
print('ok')
