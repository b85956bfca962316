"""Times representative conv_igemm launches of the B=64 forward (CUDA events, 10 reps).  With env
B200IR_DBG_SKIP_EPI=1 the epilogue only recycles the accumulators: that run gives the main-loop floor of each layer.
Usage: python tools/time_conv.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, dev = 64, 'cuda'
LAYERS = [  # (H, W, cin, cout, kind)
    (128, 384, 32, 32, 'plain'), (128, 384, 64, 32, 'res2'), (128, 384, 32, 64, 'plain'), (128, 384, 64, 64, 'mod'),
    (128, 384, 64, 64, 'modrgb'),
    (64, 192, 64, 64, 'plain'), (64, 192, 256, 64, 'res2'), (64, 192, 64, 128, 'plain'), (64, 192, 128, 128, 'modrgb'),
    (32, 96, 256, 256, 'plain'), (32, 96, 256, 512, 'plain'), (32, 96, 512, 512, 'modrgb'),
    (16, 48, 256, 256, 'plain'), (16, 48, 512, 512, 'modrgb'), (8, 24, 256, 256, 'plain'), (4, 12, 256, 256, 'plain'),
]


def build(H, W, cin, cout, kind):
    x = torch.randn(B, H, W, cin, device=dev).half()
    w = (torch.randn(cout, 9 * cin, device=dev) / math.sqrt(9 * cin)).half()
    bias = torch.zeros(cout, device=dev)
    kw = dict(bias=bias, act=True)
    keep = [x, w, bias]
    out = torch.empty(B, H, W, cout, device=dev, dtype=torch.float16)
    if kind == 'res2':
        lo = torch.randn(B, H // 2, W // 2, cout, device=dev).half()
        kw.update(res=lo, res_mode=2, res_strides=(cout, (W // 2) * cout, (H // 2) * (W // 2) * cout),
                  res_wh=(W // 2, H // 2), res_scale=0.7071)
    if kind in ('mod', 'modrgb'):
        kw.update(demod=torch.ones(B, cout, device=dev), noise=torch.randn(B, 1, H, W, device=dev),
                  noise_gain=torch.zeros(1, device=dev), noise_strides=(H * W, W))
    if kind == 'modrgb':
        kw.update(out_scale=torch.ones(B, cout, device=dev))
    if os.environ.get('NO_ROW'):
        kw.update(row_mode=0)
    op = ops.conv_same(x, w, out, 3, **kw)
    if kind == 'modrgb':
        keep.append(op.attach_rgb(torch.randn(B, 3, cout, device=dev), (H, W)))
    return op, keep


def main():
    for (H, W, cin, cout, kind) in LAYERS[:int(os.environ.get('NLAYERS', '99'))]:
        op, keep = build(H, W, cin, cout, kind)
        op()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * H * W * cout * 9 * cin
        d = op.desc
        print(f'{H:4d}x{W:<4d} {cin:4d}->{cout:<4d} {kind:7s} tile=({d.tile_b},{d.tile_h},{d.tile_w}) bn={d.block_n} '
              f'row={d.row_mode}: {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TF/s', flush=True)
        del op, keep


if __name__ == '__main__':
    main()
