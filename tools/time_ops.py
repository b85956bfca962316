"""Per-launch timing of one forward step (CUDA events, each op run `reps` times back to back after a warm-up).
Prints achieved TFLOP/s (executed MMA work) for conv_igemm launches and GB/s (tensor bytes) for the others.
Usage: python tools/time_ops.py [B] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.ops import ConvOp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
eng = net.engine()
eng.use_graphs = False
x = (torch.rand(B, 3, H, W) * 2 - 1).cuda()
net(x, return_rgb=True, randomize_noise=False)
plan = eng.plan(B)
torch.cuda.synchronize()


def label(st):
    if isinstance(st, ConvOp):
        d = st.desc
        return (f'conv taps={d.num_taps} {d.cin}->{d.cout} M=({d.m_b},{d.m_h},{d.m_w}) tile=({d.tile_b},{d.tile_h},'
                f'{d.tile_w}) bn={d.block_n} ep[{"b" if d.bias else ""}{"d" if d.demod else ""}{"n" if d.noise else ""}'
                f'{"a" if d.act else ""}r{d.res_mode}]')
    if hasattr(st, 'nbytes'):
        return st.name
    if not hasattr(st, '__code__'):
        return type(st).__name__
    names = [n for n in st.__code__.co_names if n not in ('ops', 'self', 'shape')]
    t = [v for v in (st.__defaults__ or ()) if torch.is_tensor(v)]
    return f'{names[0] if names else "?"} ' + ' '.join(str(tuple(v.shape)) for v in t[:2])


rows = []
tot = 0.0
for st in plan.steps:
    if isinstance(st, tuple):
        st = plan.rgb_steps[st[1]]
    st()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        st()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    extra = ''
    if isinstance(st, ConvOp):
        d = st.desc
        fl = 2.0 * d.m_b * d.m_h * d.m_w * d.cout * d.num_taps * d.cin
        by = 2.0 * (d.m_b * d.m_h * d.m_w * (d.cin + d.cout))
        extra = f'{fl / ms / 1e9:8.1f} TF/s  {by / ms / 1e6:8.1f} GB/s(in+out)'
    else:
        by = getattr(st, 'nbytes', 0)
        extra = f'{"":8s}       {by / ms / 1e6:8.1f} GB/s(algorithmic)'
    rows.append((ms, label(st), extra))
    print(f'{ms * 1e3:9.1f} us  {extra}  {label(st)}')
print(f'total {tot:.3f} ms for B={B} -> {B / tot * 1e3:.0f} crops/s (serialised, no overlap)')
