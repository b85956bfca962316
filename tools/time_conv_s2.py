"""Experiment: the 3x3 stride-2 conv over the FIR-smoothed buffer with (a) the four phase views striding two pixels
(current layout) and (b) the same views over a column-deinterleaved buffer (phase pixels contiguous).  Timing only.
Usage: python tools/time_conv_s2.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402
from image_restoration_b200.ops import ConvOp, View  # noqa: E402

B, dev = 64, 'cuda'


def s2_deint(p, h, w, weight, out, **kw):
    b, hp, wp, c = p.shape
    cout = weight.shape[0]
    half = wp // 2
    views = [View(p.data_ptr() + 2 * (py * wp + px * half) * c, c, half, hp // 2, b, c, 2 * wp * c, hp * wp * c)
             for py in range(2) for px in range(2)]
    taps = [((kh % 2) * 2 + (kw_ % 2), kw_ // 2, kh // 2) for kh in range(3) for kw_ in range(3)]
    oh, ow = h // 2, w // 2
    return ConvOp(views, weight, c, cout, taps, (ow, oh, b), out, (cout, ow * cout, oh * ow * cout), **kw)


def timeit(op):
    op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


for (H, W, cin, cout) in [(128, 384, 32, 64), (64, 192, 64, 256), (128, 384, 128, 256), (32, 96, 256, 256), (16, 48, 256, 256), (8, 24, 256, 256)]:
    p = torch.randn(B, H + 2, W + 2, cin, device=dev).half()
    w = (torch.randn(cout, 9 * cin, device=dev) / math.sqrt(9 * cin)).half()
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(B, H // 2, W // 2, cout, device=dev, dtype=torch.float16)
    ops._S2_FOLD = False
    a = ops.conv3x3_s2(p, H, W, w, out, bias=bias, act=True)
    b = s2_deint(p, H, W, w, out, bias=bias, act=True)
    out2 = torch.empty_like(out)
    c = ops.conv3x3_s2_folded(p, H, W, ops.s2_fold_weight(w, cin), out2, bias=bias, act=True)
    a()
    c()
    torch.cuda.synchronize()
    err = (out.float() - out2.float()).abs().max().item()
    print(f'{H}x{W} {cin}->{cout}: strided views {timeit(a):.1f} us   deinterleaved {timeit(b):.1f} us   pixel-folded '
          f'{timeit(c):.1f} us (max diff {err:.2e})   tile={a.desc.tile_w, a.desc.tile_h, a.desc.tile_b} / '
          f'{c.desc.tile_w, c.desc.tile_h, c.desc.tile_b}')
