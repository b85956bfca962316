"""Per-launch table of one GFPGANTrainer.optimize_parameters step: every C-ABI call is bracketed by CUDA events on the launching
stream and aggregated by (entry point, shape); conv / weight-gradient launches also get their executed TFLOP/s.
Usage: python tools/trace_train_step.py [batch] [fix_decoder 0|1] > gpurun_out/trace_train.txt"""
import collections
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from bench import H, NET_KW, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR, _lib, ops, train  # noqa: E402
from image_restoration_b200.disc import StyleGAN2Discriminator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fix = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
RECORDS = []
LABEL = [None]


def _val(a):
    if isinstance(a, (int, float)):
        return a
    if isinstance(a, C.c_void_p) or a is None or hasattr(a, '_obj'):
        return None
    return getattr(a, 'value', None)


class Proxy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith('b200ir_') or name in ('b200ir_last_error', 'b200ir_launch_count', 'b200ir_conv_plan_create',
                                                      'b200ir_conv_plan_destroy'):
            return fn

        def call(*args):
            ints = [v for v in (_val(a) for a in args) if isinstance(v, int) and 0 <= v < 10 ** 7]
            label, flops = name[7:] + ' ' + 'x'.join(str(v) for v in ints[:6]), 0.0
            if name == 'b200ir_conv_plan_launch' and LABEL[0] is not None:
                label, flops = LABEL[0]
            elif name == 'b200ir_conv_wgrad':
                b, h, w, cin, cout = [_val(a) for a in args[3:8]]
                flops = 2.0 * b * h * w * 9 * cin * cout
                label = f'conv_wgrad {cin}->{cout} @{h}x{w}'
            elif name == 'b200ir_conv_wgrad_view':
                v = args[0]._obj
                b, h, w, cout, mask = [_val(a) for a in args[3:8]]
                flops = 2.0 * b * h * w * bin(mask).count('1') * v.c * cout
                label = f'conv_wgrad_view {v.c}->{cout} @{h}x{w} taps={bin(mask).count("1")}'
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args)
            e1.record()
            RECORDS.append((label, flops, e0, e1))
            return r
        return call


_orig_call = ops.ConvOp.__call__


def _conv_call(self):
    d = self.desc
    fl = 2.0 * d.m_b * d.m_h * d.m_w * d.cout * d.num_taps * d.cin
    LABEL[0] = (f'conv taps={d.num_taps} {d.cin}->{d.cout} M=({d.m_b},{d.m_h},{d.m_w}) row={d.row_mode} '
                f'ep[{"b" if d.bias else ""}{"d" if d.demod else ""}{"n" if d.noise else ""}{"a" if d.act else ""}r{d.res_mode}]', fl)
    _orig_call(self)
    LABEL[0] = None


torch.manual_seed(0)
kw = dict(NET_KW, fix_decoder=fix)
net = GFPGANv1OCR(**kw).cuda().train()
netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1).cuda()
tr = train.GFPGANTrainer(net, netd)
gt = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
lq = (gt + 0.1 * torch.randn_like(gt)).clamp(-1, 1)
for it in range(2):
    tr.feed_data(lq, gt)
    tr.optimize_parameters(it + 1)
torch.cuda.synchronize()
real = _lib.lib()
_lib._lib = Proxy(real)
ops.ConvOp.__call__ = _conv_call
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
tr.feed_data(lq, gt)
tr.optimize_parameters(3)
t1.record()
torch.cuda.synchronize()
_lib._lib = real
agg = collections.OrderedDict()
for label, flops, e0, e1 in RECORDS:
    a = agg.setdefault(label, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += flops
tot = sum(a[1] for a in agg.values())
print(f'step {t0.elapsed_time(t1):.1f} ms, inside C-ABI calls {tot:.1f} ms, {len(RECORDS)} launches, B={B}, fix_decoder={fix}')
for label, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = f'{fl / ms / 1e9:7.0f} TF/s' if fl else ' ' * 12
    print(f'{ms:8.2f} ms {100 * ms / tot:5.1f}% {n:4d}x {tf}  {label}')
