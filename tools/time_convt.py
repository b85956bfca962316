"""Merged transposed conv (one GEMM) vs the four phase launches at the decoder levels of the B=64 forward.
B200IR_DBG_SKIP_EPI=1 gives the main-loop floor."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, dev = 64, 'cuda'


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (h, w, cin, cout) in [(64, 192, 128, 64), (32, 96, 512, 128), (16, 48, 512, 512), (8, 24, 512, 512)]:
    x = torch.randn(B, h, w, cin, device=dev).half()
    wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
    demod = torch.ones(B, cout, device=dev)
    raw = torch.zeros(B, 2 * h + 2, 2 * w + 2, cout, device=dev, dtype=torch.float16)
    merged = ops.convt_s2_merged(x, ops.convt_merged_weight(wt, 1.0), raw, demod)
    ws = wt.half()
    phases = [ops.convt_s2_phase(x, torch.cat([ws[:, :, kh, kw] for kh, kw in ops.convt_phase_taps(py, px)], 1).contiguous(),
                                 py, px, raw, demod=demod) for py, px in ops.CONVT_PHASES]

    def run_phases():
        for p in phases:
            p()
    d = merged.desc
    fl = 2.0 * B * h * w * 9 * cin * cout
    t_m, t_p = timeit(merged), timeit(run_phases)
    print(f'{h}x{w} {cin}->{cout}: merged tile=({d.tile_b},{d.tile_h},{d.tile_w}) bn={d.block_n} {t_m:7.1f} us '
          f'({fl / t_m / 1e6:6.0f} TF/s alg)   4 phases {t_p:7.1f} us ({fl / t_p / 1e6:6.0f} TF/s alg)', flush=True)
