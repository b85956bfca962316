"""Experiment: tile shape of the merged transposed conv 128 -> 4x64 @64x192 (M = 65 x 193 per image).  Usage: python tools/time_convt_tiles.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402
from image_restoration_b200.ops import ConvOp, nhwc_view  # noqa: E402

B, dev = 64, 'cuda'
for (h, w, cin, cout) in [(64, 192, 128, 64), (32, 96, 512, 128)]:
    x = torch.randn(B, h, w, cin, device=dev).half()
    wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
    w_big = ops.convt_merged_weight(wt, 1.0)
    demod = torch.ones(B, cout, device=dev)
    raw = torch.zeros(B, 2 * h + 2, 2 * w + 2, cout, device=dev, dtype=torch.float16)
    ref = None
    for tile in [None, (8, 2, 8), (16, 2, 4), (32, 2, 2), (64, 2, 1), (16, 8, 1), (32, 4, 1), (8, 4, 4), (16, 4, 2), (8, 16, 1), (128, 1, 1)]:
        taps = [(0, -tx, -ty) for ty in range(2) for tx in range(2)]
        bn = min(256, 4 * cout)
        masks = []
        for j in range(4 * cout // bn):
            m = 0
            for ph in range(j * bn // cout, ((j + 1) * bn - 1) // cout + 1):
                py, px = ph // 2, ph % 2
                for ty in range(2):
                    for tx in range(2):
                        if py + 2 * ty <= 2 and px + 2 * tx <= 2:
                            m |= 1 << (ty * 2 + tx)
            masks.append(m)
        t = tile or ops.pick_tile(w + 1, h + 1, B, min_w=8, max_b=max(1, 2560 // bn))
        _, rh, rw, _ = raw.shape
        raw.zero_()
        op = ConvOp([nhwc_view(x)], w_big, cin, 4 * cout, taps, (w + 1, h + 1, B), raw, (cout, rw * cout, rh * rw * cout),
                    demod=demod, block_n=bn, tile=t, ps_r=2, ps_c=cout, demod_c=cout, tap_mask=masks)
        try:
            op()
            torch.cuda.synchronize()
        except Exception as e:
            print(f'{h}x{w} {cin}->{cout} tile={t}: {str(e)[:80]}')
            continue
        if ref is None:
            ref = raw.clone()
        same = torch.equal(raw[:, :2 * h + 1, :2 * w + 1], ref[:, :2 * h + 1, :2 * w + 1])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        tw, th, tb = t
        waste = (-(-(w + 1) // tw) * tw) * (-(-(h + 1) // th) * th) * (-(-B // tb) * tb) / ((w + 1) * (h + 1) * B)
        print(f'{h}x{w} {cin}->{cout} tile(w,h,b)={t}: {us:7.1f} us  padded/valid {waste:.3f}  same={same}')
