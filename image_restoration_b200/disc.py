"""StyleGAN2Discriminator (Car_Plate-Restoration/basicsr/archs/stylegan2_arch.py:735-805), the `network_d` of the training
YAMLs (training_config/*.yml), forward pass on the B200 kernels.  First piece of the training step:
the discriminator scores that the GAN loss of GFPGANModel.optimize_parameters reads.  Forward only -- no autograd.

Same constructor keywords, parameter names / shapes and seeded initialisation as the reference class, so its
checkpoints (`net_d_*.pth`) load with strict=True.  The modules only hold parameters; the arithmetic runs in
DiscEngine: ConvLayer(3, C, 1) = first_conv; ResBlock = the same launches as the U-Net encoder of the restoration
network (3x3 conv, FIR + stride-2 3x3 conv with the FIR + 1x1 skip as a fused residual); minibatch stddev
(b200ir_minibatch_stddev); final_conv over the concatenated tensor (channels padded 513 -> 528 with zero weights);
the two EqualLinear layers as 1x1 convs over a [B,1,1,K] view.
"""
import math
import weakref

import torch
from torch import nn

from . import _lib, ops
from .arch import _Linear, _ResBlock, _conv_layer
from .registry import ARCH_REGISTRY, USING_BASICSR_REGISTRY

_ENGINES = weakref.WeakKeyDictionary()
F16, F32 = torch.float16, torch.float32


class StyleGAN2Discriminator(nn.Module):
    def __init__(self, input_width=256, input_height=256, channel_multiplier=2, resample_kernel=(1, 3, 3, 1),
                 stddev_group=4, narrow=1):
        super().__init__()
        if tuple(resample_kernel) != (1, 3, 3, 1):
            raise ValueError('the FIR kernels implement resample_kernel=(1, 3, 3, 1) (every shipped config)')
        out_size = min(input_width, input_height)
        ch = {4: int(512 * narrow), 8: int(512 * narrow), 16: int(512 * narrow), 32: int(512 * narrow),
              64: int(256 * channel_multiplier * narrow), 128: int(128 * channel_multiplier * narrow),
              256: int(64 * channel_multiplier * narrow), 512: int(32 * channel_multiplier * narrow),
              1024: int(16 * channel_multiplier * narrow)}
        log_size = int(math.log(out_size, 2))
        body = [_conv_layer(3, ch[out_size], 1)]
        cin = ch[out_size]
        for i in range(log_size, 2, -1):
            cout = ch[2 ** (i - 1)]
            body.append(_ResBlock(cin, cout))
            cin = cout
        self.conv_body = nn.Sequential(*body)
        self.final_conv = _conv_layer(cin + 1, ch[4], 3)
        ratio = int(input_width / input_height)
        self.final_linear = nn.Sequential(_Linear(ch[4] * 4 * 4 * ratio, ch[4]), _Linear(ch[4], 1))
        self.stddev_group, self.stddev_feat = stddev_group, 1
        self.input_width, self.input_height, self.log_size = input_width, input_height, log_size

    def engine(self):
        eng = _ENGINES.get(self)
        if eng is None:
            eng = DiscEngine(self)
            _ENGINES[self] = eng
        return eng

    def _apply(self, fn, *a, **kw):
        _ENGINES.pop(self, None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        _ENGINES.pop(self, None)
        return super().load_state_dict(*a, **kw)

    def forward(self, x):
        """x: (B,3,H,W) float CUDA tensor -> (B,1) scores, as stylegan2_arch.py:788-805."""
        if not x.is_cuda:
            raise RuntimeError('image_restoration_b200.StyleGAN2Discriminator runs on a CUDA B200 only (no CPU path)')
        return self.engine().forward(x)


class DiscEngine:
    """Packed weights + per-batch-size launch lists (eager launches on the caller's stream)."""

    def __init__(self, net):
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('image_restoration_b200 needs the module on a CUDA B200 (no CPU path)')
        _lib.lib()
        self.net, self.dev = net, dev
        sd = {k: v.detach().to(dev) for k, v in net.state_dict().items()}

        def eq(w, cin_pad=None):
            w = w.float()
            scale = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])       # EqualConv2d (stylegan2_arch.py:639-648)
            if cin_pad is not None and cin_pad > w.shape[1]:
                w = torch.cat([w, w.new_zeros(w.shape[0], cin_pad - w.shape[1], w.shape[2], w.shape[3])], 1)
            co, ci, kh, kw = w.shape
            return (w * scale).permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(F16)

        f32 = lambda k: sd[k].float().contiguous()  # noqa: E731
        w0 = sd['conv_body.0.0.weight'].float()
        self.first_w = (w0[:, :, 0, 0] / math.sqrt(3.0)).contiguous()
        self.first_b = f32('conv_body.0.1.bias')
        self.blocks = []
        for i in range(1, len(net.conv_body)):
            p = f'conv_body.{i}'
            self.blocks.append(dict(w1=eq(sd[f'{p}.conv1.0.weight']), b1=f32(f'{p}.conv1.1.bias'),
                                    w2=eq(sd[f'{p}.conv2.1.weight']), b2=f32(f'{p}.conv2.2.bias'),
                                    ws=eq(sd[f'{p}.skip.1.weight'])))
        wf = sd['final_conv.0.weight']
        self.c_last = wf.shape[1] - 1
        self.c_pad = (wf.shape[1] + 15) // 16 * 16
        self.final_w = eq(wf, self.c_pad)
        self.final_b = f32('final_conv.1.bias')
        # EqualLinear (stylegan2_arch.py:134-175, lr_mul = 1); the first one reads the NCHW-flattened feature map
        w1 = sd['final_linear.0.weight'].float()
        n1, k1 = w1.shape
        c4 = self.final_w.shape[0]
        P = k1 // c4
        self.lin1_w = (w1.view(n1, c4, P).permute(0, 2, 1).reshape(n1, k1) / math.sqrt(k1)).contiguous().to(F16)
        self.lin1_b = f32('final_linear.0.bias')
        w2 = sd['final_linear.1.weight'].float()
        self.lin2_w = torch.cat([w2 / math.sqrt(w2.shape[1]), w2.new_zeros(15, w2.shape[1])], 0).contiguous().to(F16)
        self.lin2_b = torch.cat([f32('final_linear.1.bias'), torch.zeros(15, device=dev)]).contiguous()
        self.plans = {}

    def plan(self, B):
        pl = self.plans.get(B)
        if pl is not None:
            return pl
        net, dev = self.net, self.dev
        group = min(B, net.stddev_group)
        if B % group:
            raise ValueError(f'batch {B} is not divisible by the stddev group {group} (the reference fails here too)')
        H, W = net.input_height, net.input_width
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=F16)  # noqa: E731
        z16 = lambda *s: torch.zeros(*s, device=dev, dtype=F16)  # noqa: E731
        steps = []
        x_in = torch.empty(B, 3, H, W, device=dev, dtype=F32)
        feat = e16(B, H, W, self.first_w.shape[0])
        steps.append(lambda o=feat: ops.first_conv(x_in, self.first_w, self.first_b, o))
        h, w = H, W
        for d in self.blocks:                                    # ResBlock (stylegan2_arch.py:706-732)
            cin, cout = d['w1'].shape[0], d['w2'].shape[0]
            t1 = e16(B, h, w, cin)
            steps.append(ops.conv_same(feat, d['w1'], t1, 3, bias=d['b1'], act=True))
            p = z16(B, h + 2, w + 2, cin)
            steps.append(lambda a=t1, o=p: ops.fir_pad22(a, o))
            sk_in = e16(B, h // 2, w // 2, cin)
            steps.append(lambda a=feat, o=sk_in: ops.fir_down2(a, o))
            sk = e16(B, h // 2, w // 2, cout)
            steps.append(ops.conv_same(sk_in, d['ws'], sk, 1))
            oh, ow = h // 2, w // 2
            nxt = e16(B, oh, ow, cout)
            steps.append(ops.conv3x3_s2(p, h, w, d['w2'], nxt, bias=d['b2'], act=True, res=sk, res_mode=1,
                                        res_strides=(cout, ow * cout, oh * ow * cout), res_wh=(ow, oh),
                                        res_scale=ops.INV_SQRT2))
            feat, h, w = nxt, oh, ow
        cat = e16(B, h, w, self.c_pad)
        s_buf = torch.empty(B // group, device=dev, dtype=F32)
        lib = _lib.lib()

        def mbstd(a=feat, o=cat):
            _lib.check(lib.b200ir_minibatch_stddev(ops._ptr(a), ops._ptr(s_buf), ops._ptr(o), B, h * w, self.c_last,
                                                   self.c_pad, group, ops._stream()), 'minibatch_stddev')
        steps.append(mbstd)
        c4 = self.final_w.shape[0]
        g = e16(B, h, w, c4)
        steps.append(ops.conv_same(cat, self.final_w, g, 3, bias=self.final_b, act=True))
        k1 = h * w * c4
        hid = e16(B, c4)
        v1 = ops.View(g.data_ptr(), k1, 1, 1, B, k1, k1, k1)
        steps.append(ops.ConvOp([v1], self.lin1_w, k1, c4, [(0, 0, 0)], (1, 1, B), hid, (c4, c4, c4), tile=(1, 1, 128),
                                bias=self.lin1_b, act=True, block_n=64))     # fused_lrelu(out, bias)
        score = torch.empty(B, 16, device=dev, dtype=F32)
        steps.append(ops.linear_as_conv(hid, self.lin2_w, score, bias=self.lin2_b, block_n=16))
        pl = dict(x_in=x_in, steps=steps, score=score, keep=(feat, cat, g, hid, s_buf))
        self.plans[B] = pl
        return pl

    @torch.no_grad()
    def forward(self, x):
        B = x.shape[0]
        net = self.net
        if tuple(x.shape[1:]) != (3, net.input_height, net.input_width):
            raise ValueError(f'expected (B,3,{net.input_height},{net.input_width}), got {tuple(x.shape)}')
        with torch.cuda.device(self.dev):
            pl = self.plan(B)
            pl['x_in'].copy_(x)
            for st in pl['steps']:
                st()
            return pl['score'][:, :1].clone().to(x.dtype)


if 'StyleGAN2Discriminator_B200' not in ARCH_REGISTRY:
    ARCH_REGISTRY.register(type('StyleGAN2Discriminator_B200', (StyleGAN2Discriminator,),
                                {'__doc__': StyleGAN2Discriminator.__doc__}))
if not USING_BASICSR_REGISTRY and 'StyleGAN2Discriminator' not in ARCH_REGISTRY:
    ARCH_REGISTRY.register(StyleGAN2Discriminator)
