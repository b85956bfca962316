"""Experiment: does running the HBM-heavy top level of the StyleGAN2 decoder (transposed conv 128 -> 4x64 @64x192 ->
upfir_act -> StyleConv 64 -> 64 @128x384 with fused ToRGB, nothing stored) in sub-batches keep the intermediate tensors in the
126 MB L2?  Times the three-launch chain for B = 64 as one pass and as sub-batches of 32 / 16 / 8 crops that reuse the same
scratch buffers (CUDA graph replays, CUDA events).  Usage: python tools/time_l2_subbatch.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops  # noqa: E402

B, dev = 64, 'cuda'
h, w, cin, cout = 64, 192, 128, 64
h2, w2 = 2 * h, 2 * w
torch.manual_seed(0)
xs = torch.randn(B, h, w, cin, device=dev).half()
wt = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
w_big = ops.convt_merged_weight(wt, 1.0)
demod1 = torch.rand(B, cout, device=dev) + 0.5
demod2 = torch.rand(B, cout, device=dev) + 0.5
w2c = (torch.randn(cout, 9 * cout, device=dev) / math.sqrt(9 * cout)).half()
bias = torch.zeros(cout, device=dev)
gain = torch.zeros(1, device=dev)
n1 = torch.randn(B, 1, h2, w2, device=dev)
n2 = torch.randn(B, 1, h2, w2, device=dev)
c_sft = cout // 2
sc = torch.randn(B, h2, w2, c_sft, device=dev).half()
sh = torch.randn(B, h2, w2, c_sft, device=dev).half()
s_next = torch.ones(B, cout, device=dev)
wm = torch.randn(B, 3, cout, device=dev)


def chain(sb):
    """launch list of the chain in sub-batches of sb crops; scratch buffers are shared by all sub-batches when sb < B"""
    raw = torch.zeros(sb, h2 + 2, w2 + 2, cout, device=dev, dtype=torch.float16)
    xs2 = torch.empty(sb, h2, w2, cout, device=dev, dtype=torch.float16)
    launches, keep = [], [raw, xs2]
    for b0 in range(0, B, sb):
        sl = slice(b0, b0 + sb)
        launches.append(ops.convt_s2_merged(xs[sl], w_big, raw, demod1[sl].contiguous()))
        launches.append(lambda r=raw, o=xs2, n=n1[sl], a=sc[sl], b_=sh[sl], sn=s_next[sl].contiguous():
                        ops.upfir_act(r, o, n, h2 * w2, gain, bias, a, b_, c_sft, sn))
        op = ops.conv_same(xs2, w2c, None, 3, bias=bias, demod=demod2[sl].contiguous(), noise=n2[sl], noise_gain=gain,
                           noise_strides=(h2 * w2, w2), act=True)
        keep.append(op.attach_rgb(wm[sl].contiguous(), (h2, w2), no_store=True))
        launches.append(op)
    return launches, keep


for sb in (64, 32, 16, 8, 64, 8):
    launches, keep = chain(sb)
    for f in launches:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in launches:
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'sub-batch {sb:3d}: {len(launches):3d} launches, {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per 64 crops', flush=True)
    del g, launches, keep
