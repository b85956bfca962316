"""Times selected launches of the B=64 forward plan alone (CUDA events, 20 reps each); with B200IR_DBG_SKIP_EPI=1 in the
environment the epilogues only recycle the accumulators, which gives each layer's main-loop floor.
Usage: python tools/time_plan_ops.py [min_us]   (all conv launches slower than min_us, default 60)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.ops import ConvOp  # noqa: E402

B = 64
min_us = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
eng = net.engine()
eng.use_graphs = False
x = (torch.rand(B, 3, H, W) * 2 - 1).cuda()
net(x, return_rgb=False, randomize_noise=False)
plan = eng.plan(B)
torch.cuda.synchronize()
tot = 0.0
for st in plan.steps:
    if not isinstance(st, ConvOp):
        continue
    d = st.desc
    st()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        st()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    tot += us
    if us >= min_us:
        fl = 2.0 * d.m_b * d.m_h * d.m_w * d.cout * d.num_taps * d.cin
        print(f'{us:8.1f} us {fl / us / 1e6:7.0f} TF/s  taps={d.num_taps} {d.cin}->{d.cout} M=({d.m_b},{d.m_h},{d.m_w}) '
              f'tile=({d.tile_b},{d.tile_h},{d.tile_w}) bn={d.block_n} row={d.row_mode} '
              f'ep[{"b" if d.bias else ""}{"d" if d.demod else ""}{"n" if d.noise else ""}{"a" if d.act else ""}r{d.res_mode}'
              f'{"R" if d.rgb_w else ""}{"U" if d.corr_top else ""}]')
print(f'total conv {tot:.0f} us, skip_epi={os.environ.get("B200IR_DBG_SKIP_EPI", "0")}')
