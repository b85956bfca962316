"""Times the streaming FIR kernels at the five decoder / encoder shapes of the B=64 forward under tuning knobs
(env B200IR_FIR_XP, B200IR_FIR_R).  Usage: python tools/time_fir.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B = 64
dev = 'cuda'


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run():
    res = []
    for (h2, w2, C) in [(128, 384, 64), (64, 192, 128), (32, 96, 512), (16, 48, 512)]:
        raw = torch.randn(B, h2 + 2, w2 + 2, C, device=dev).half()
        out = torch.empty(B, h2, w2, C, device=dev, dtype=torch.float16)
        noise = torch.randn(B, 1, h2, w2, device=dev)
        gain = torch.zeros(1, device=dev)
        bias = torch.zeros(C, device=dev)
        sc = torch.randn(B, h2, w2, C // 2, device=dev).half()
        sh = torch.randn(B, h2, w2, C // 2, device=dev).half()
        sn = torch.ones(B, C, device=dev)
        ms = timeit(lambda: ops.upfir_act(raw, out, noise, h2 * w2, gain, bias, sc, sh, C // 2, sn))
        by = (raw.numel() + out.numel() + sc.numel() * 2) * 2
        res.append(f'upfir {h2}x{w2}x{C}: {ms * 1e3:7.1f} us {by / ms / 1e6:7.0f} GB/s')
    for (H, W, C) in [(128, 384, 32), (64, 192, 64), (32, 96, 256)]:
        x = torch.randn(B, H, W, C, device=dev).half()
        p = torch.zeros(B, H + 2, W + 2, C, device=dev, dtype=torch.float16)
        ms = timeit(lambda: ops.fir_pad22(x, p))
        by = (x.numel() + p.numel()) * 2
        res.append(f'pad22 {H}x{W}x{C}: {ms * 1e3:7.1f} us {by / ms / 1e6:7.0f} GB/s')
    return res


for nt in ('0', '256', '384'):
    for xp in ('0', '1', '2'):
        os.environ['B200IR_FIR_XP'] = xp
        os.environ['B200IR_FIR_NT'] = nt
        print(f'--- NT={nt} XP={xp}')
        for line in run():
            print('   ', line)
