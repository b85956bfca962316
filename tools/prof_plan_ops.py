"""Runs selected launches of the B=64 forward plan alone (for ncu captures with --profile-from-start off: the selected
launches sit between cudaProfilerStart / Stop).  Selectors: convt128 (merged transposed conv 128->256 @64x192), upfold (folded
ConvUpLayer 256->4x64), mod128 (StyleConv 128->128 @64x192), s2_32 (stride-2 conv 32->64), row32 (32->32 row kernel),
fir (all fir_pad22 launches).  Usage: python tools/prof_plan_ops.py convt128 upfold"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.ops import ConvOp  # noqa: E402

B = 64
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
eng = net.engine()
eng.use_graphs = False
x = (torch.rand(B, 3, H, W) * 2 - 1).cuda()
net(x, return_rgb=False, randomize_noise=False)
plan = eng.plan(B)
torch.cuda.synchronize()


def pick(sel):
    out = []
    for st in plan.steps:
        if isinstance(st, ConvOp):
            d = st.desc
            key = (d.num_taps, d.cin, d.cout, d.m_h, d.m_w)
            if sel == 'convt128' and d.num_taps == 4 and d.cin == 128 and d.cout == 256:
                out.append(st)
            elif sel == 'upfold' and d.corr_top:
                out.append(st)
            elif sel == 'mod128' and key == (9, 128, 128, 64, 192):
                out.append(st)
            elif sel == 's2_32' and d.num_views == 2 and d.cin == 64 and d.cout == 64 and d.num_taps == 6:
                out.append(st)
            elif sel == 'row32' and key == (9, 32, 32, 128, 384) and d.row_mode:
                out.append(st)
        elif sel == 'fir' and getattr(st, 'name', '') == 'fir_pad22':
            out.append(st)
    return out


ops_sel = []
for sel in sys.argv[1:]:
    found = pick(sel)
    print(sel, len(found), 'launches')
    ops_sel += found[:2]
for op in ops_sel:
    op()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for op in ops_sel:
    for _ in range(2):
        op()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok')
