"""Drop-in MSRResNet / EDSR / RCAN / RRDBNet for B200: the networks the reference's `options/*.yml` select from ARCH_REGISTRY.

Mirrors `basicsr/archs/srresnet_arch.py:8-68` (MSRResNet), `edsr_arch.py:8-72` (EDSR), `rcan_arch.py:8-135` (RCAN),
`rrdbnet_arch.py:9-123` (RRDBNet: dense concatenation = channel-offset outputs into one NHWC buffer, nearest x2, pixel_unshuffle) and
the blocks of `arch_util.py` they use (`ResidualBlockNoBN :66-93`, `Upsample :96-109`, `default_init_weights :12-43`):
same constructor keywords, same parameter names / shapes (checkpoints load unchanged), same `forward(x)`.
The modules only hold parameters; every arithmetic op runs in libb200ir.so:

  * all 3x3 convolutions -> `b200ir_conv_igemm` (tcgen05 implicit GEMM) with ReLU / LeakyReLU(0.1), the
    `identity + out * res_scale` merge and `nn.PixelShuffle` fused into the epilogue (act == 2, res_mul, ps_r);
  * mean shift + NCHW fp32 -> NHWC fp16 (3 -> 16 zero-padded channels), output assembly with the bilinear base image,
    RCAN channel attention (global average pool, squeeze/excite 1x1s, sigmoid, scale + residual) -> sr_ops.cu.

Activations are NHWC fp16, accumulation fp32; the last conv writes fp32.  No CPU / torch fallback.
"""
import math
import weakref

import torch
from torch import nn
from torch.nn import init

from . import ops
from .registry import ARCH_REGISTRY, USING_BASICSR_REGISTRY

F16, F32 = torch.float16, torch.float32
_ENGINES = weakref.WeakKeyDictionary()


# ------------------------------------------------------------------------------------------ parameter containers
@torch.no_grad()
def _default_init_weights(modules, scale=1.0):
    """arch_util.default_init_weights (:12-43) for conv layers: kaiming_normal_ * scale, zero bias."""
    for module in modules:
        for m in module.modules():
            if isinstance(m, nn.Conv2d):
                init.kaiming_normal_(m.weight)
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.fill_(0)


class _ResidualBlockNoBN(nn.Module):
    """Parameters of arch_util.ResidualBlockNoBN (:66-93): conv1 -> ReLU -> conv2, identity + out * res_scale."""

    def __init__(self, num_feat=64, res_scale=1, pytorch_init=False):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        if not pytorch_init:
            _default_init_weights([self.conv1, self.conv2], 0.1)


class _Marker(nn.Module):
    """Parameter-free stand-in (PixelShuffle / ReLU / pooling / Sigmoid) that keeps nn.Sequential indices identical."""


def _upsample(scale, num_feat):
    """arch_util.Upsample (:96-109): [conv(num_feat -> r^2 num_feat), PixelShuffle(r)] x log2(scale), or one x3 stage."""
    m = []
    if (scale & (scale - 1)) == 0:
        for _ in range(int(math.log(scale, 2))):
            m += [nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1), _Marker()]
    elif scale == 3:
        m += [nn.Conv2d(num_feat, 9 * num_feat, 3, 1, 1), _Marker()]
    else:
        raise ValueError(f'scale {scale} is not supported. Supported scales: 2^n and 3.')
    return nn.Sequential(*m)


class _ChannelAttention(nn.Module):
    def __init__(self, num_feat, squeeze_factor=16):
        super().__init__()
        self.attention = nn.Sequential(_Marker(), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0), _Marker(),
                                       nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0), _Marker())


class _RCAB(nn.Module):
    def __init__(self, num_feat, squeeze_factor=16, res_scale=1):
        super().__init__()
        self.res_scale = res_scale
        self.rcab = nn.Sequential(nn.Conv2d(num_feat, num_feat, 3, 1, 1), _Marker(), nn.Conv2d(num_feat, num_feat, 3, 1, 1),
                                  _ChannelAttention(num_feat, squeeze_factor))


class _ResidualGroup(nn.Module):
    def __init__(self, num_feat, num_block, squeeze_factor=16, res_scale=1):
        super().__init__()
        self.residual_group = nn.Sequential(*[_RCAB(num_feat, squeeze_factor, res_scale) for _ in range(num_block)])
        self.conv = nn.Conv2d(num_feat, num_feat, 3, 1, 1)


class _SrBase(nn.Module):
    def engine(self):
        eng = _ENGINES.get(self)
        if eng is None or eng.stale():
            eng = SrEngine(self)
            _ENGINES[self] = eng
        return eng

    def _apply(self, fn, *a, **kw):
        _ENGINES.pop(self, None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        _ENGINES.pop(self, None)
        return super().load_state_dict(*a, **kw)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError(f'image_restoration_b200.{type(self).__name__} runs on a CUDA B200 only; there is no CPU path')
        return self.engine().forward(x)


class MSRResNet(_SrBase):
    """srresnet_arch.MSRResNet (:8-68)."""

    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4):
        super().__init__()
        self.upscale, self.num_feat, self.num_in_ch, self.num_out_ch = upscale, num_feat, num_in_ch, num_out_ch
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[_ResidualBlockNoBN(num_feat=num_feat) for _ in range(num_block)])
        if upscale in (2, 3):
            self.upconv1 = nn.Conv2d(num_feat, num_feat * upscale * upscale, 3, 1, 1)
        elif upscale == 4:
            self.upconv1 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1)
            self.upconv2 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1)
        else:
            raise ValueError(f'upscale {upscale} is not supported (x2, x3, x4)')
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        _default_init_weights([self.conv_first, self.upconv1, self.conv_hr, self.conv_last], 0.1)
        if upscale == 4:
            _default_init_weights([self.upconv2], 0.1)


class EDSR(_SrBase):
    """edsr_arch.EDSR (:8-72)."""

    def __init__(self, num_in_ch, num_out_ch, num_feat=64, num_block=16, upscale=4, res_scale=1, img_range=255.,
                 rgb_mean=(0.4488, 0.4371, 0.4040)):
        super().__init__()
        self.upscale, self.num_feat, self.num_in_ch, self.num_out_ch = upscale, num_feat, num_in_ch, num_out_ch
        self.img_range = img_range
        self.mean = torch.Tensor(rgb_mean).view(1, 3, 1, 1)     # plain attribute in the reference too (not a buffer)
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[_ResidualBlockNoBN(num_feat=num_feat, res_scale=res_scale, pytorch_init=True)
                                    for _ in range(num_block)])
        self.conv_after_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.upsample = _upsample(upscale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)


class RCAN(_SrBase):
    """rcan_arch.RCAN (:69-135)."""

    def __init__(self, num_in_ch, num_out_ch, num_feat=64, num_group=10, num_block=16, squeeze_factor=16, upscale=4,
                 res_scale=1, img_range=255., rgb_mean=(0.4488, 0.4371, 0.4040)):
        super().__init__()
        self.upscale, self.num_feat, self.num_in_ch, self.num_out_ch = upscale, num_feat, num_in_ch, num_out_ch
        self.img_range = img_range
        self.mean = torch.Tensor(rgb_mean).view(1, 3, 1, 1)
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[_ResidualGroup(num_feat, num_block, squeeze_factor, res_scale)
                                    for _ in range(num_group)])
        self.conv_after_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.upsample = _upsample(upscale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)


class _ResidualDenseBlock(nn.Module):
    """Parameters of rrdbnet_arch.ResidualDenseBlock (:9-39)."""

    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        _default_init_weights([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], 0.1)


class _RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1 = _ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = _ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = _ResidualDenseBlock(num_feat, num_grow_ch)


class RRDBNet(_SrBase):
    """rrdbnet_arch.RRDBNet (:66-123)."""

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale, self.num_feat, self.num_grow_ch = scale, num_feat, num_grow_ch
        self.num_in_ch, self.num_out_ch = num_in_ch, num_out_ch           # of the image (before pixel_unshuffle)
        self.upscale = scale
        cin = num_in_ch * (4 if scale == 2 else (16 if scale == 1 else 1))
        self.conv_first = nn.Conv2d(cin, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[_RRDB(num_feat, num_grow_ch=num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)


# ------------------------------------------------------------------------------------------ engine
def _pack(conv, cin_pad=None, cout_pad=None, ps_r=0):
    """nn.Conv2d 3x3 -> (fp16 [cout][9*cin] tap-major, fp32 bias); optional zero padding of cin / cout; for a conv
    followed by PixelShuffle(r) the output rows are reordered from (c, dy, dx) to (dy, dx, c) (module docstring)."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    co, ci, kh, kw = w.shape
    if ps_r:
        c = co // (ps_r * ps_r)
        perm = torch.arange(co, device=w.device).view(c, ps_r * ps_r).t().reshape(-1)     # new row t*c + c' <- c'*r^2 + t
        w, b = w[perm], b[perm]
    if cin_pad and cin_pad > ci:
        w = torch.cat([w, w.new_zeros(co, cin_pad - ci, kh, kw)], 1)
    if cout_pad and cout_pad > co:
        w = torch.cat([w, w.new_zeros(cout_pad - co, w.shape[1], kh, kw)], 0)
        b = torch.cat([b, b.new_zeros(cout_pad - co)])
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(F16), b.contiguous()


class _SrPlan:
    """Buffers + prepared launches for one input shape (B, H, W)."""

    def __init__(self, eng, B, H, W):
        net, dev = eng.net, eng.dev
        nf, r_total = net.num_feat, net.upscale
        steps = []
        self.steps = steps
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=F16)  # noqa: E731
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=F32)  # noqa: E731
        self.keep = []

        def conv(x, mod, out=None, *, slope=None, res=None, res_scale=1.0, res_mul=0.0, ps_r=0, cin_pad=None,
                 cout_pad=None, out_fp32=False):
            w, b = eng.packed(mod, cin_pad, cout_pad, ps_r)
            bb, h, ww, _ = x.shape
            cout = w.shape[0]
            if out is None:
                c_out = cout // (ps_r * ps_r) if ps_r else cout
                rr = ps_r or 1
                out = (e32 if out_fp32 else e16)(bb, h * rr, ww * rr, c_out)
            kw = dict(bias=b, act_slope=slope, out_fp32=out_fp32)
            if res is not None:
                c = res.shape[3]
                kw.update(res=res, res_mode=1, res_strides=(c, ww * c, h * ww * c), res_wh=(ww, h), res_scale=res_scale,
                          res_mul=res_mul)
            if ps_r:
                c_out = cout // (ps_r * ps_r)
                oh, ow = h * ps_r, ww * ps_r
                op = ops.ConvOp([ops.nhwc_view(x)], w, x.shape[3], cout, ops.taps_3x3(), (ww, h, bb), out,
                                (c_out, ow * c_out, oh * ow * c_out), block_n=c_out, ps_r=ps_r, **kw)
            else:
                op = ops.conv_same(x, w, out, 3, **kw)
            steps.append(op)
            return out

        # ---- input: mean shift (EDSR / RCAN), NCHW fp32 -> NHWC fp16 with 16 channels
        self.x_in = e32(B, net.num_in_ch, H, W)
        cin_pad = 16
        x16 = e16(B, H, W, cin_pad)
        sub = eng.mean_dev if hasattr(net, 'img_range') else None
        mul = float(net.img_range) if hasattr(net, 'img_range') else 1.0
        if not isinstance(net, RRDBNet):
            steps.append(lambda: ops.nchw_to_nhwc_pad(self.x_in, x16, sub, mul))

        if isinstance(net, RRDBNet):
            self._build_rrdb(eng, B, H, W, conv, e16, e32)
        elif isinstance(net, MSRResNet):
            # srresnet_arch.py:55-68
            feat = conv(x16, net.conv_first, slope=0.1, cin_pad=cin_pad)
            out = feat
            for blk in net.body:
                t = conv(out, blk.conv1, slope=0.0)
                out = conv(t, blk.conv2, res=out, res_scale=blk.res_scale, res_mul=1.0)
            if net.upscale == 4:
                out = conv(out, net.upconv1, slope=0.1, ps_r=2)
                out = conv(out, net.upconv2, slope=0.1, ps_r=2)
            else:
                out = conv(out, net.upconv1, slope=0.1, ps_r=net.upscale)
            out = conv(out, net.conv_hr, slope=0.1)
            y = conv(out, net.conv_last, cout_pad=16, out_fp32=True)
            self.out = e32(B, net.num_out_ch, H * r_total, W * r_total)
            steps.append(lambda: ops.sr_output(y, self.out, 1.0, None, self.x_in, r_total))
        else:
            # edsr_arch.py:59-72 / rcan_arch.py:122-135
            x = conv(x16, net.conv_first, cin_pad=cin_pad)
            out = x
            if isinstance(net, EDSR):
                for blk in net.body:
                    t = conv(out, blk.conv1, slope=0.0)
                    out = conv(t, blk.conv2, res=out, res_scale=blk.res_scale, res_mul=1.0)
            else:
                for grp in net.body:
                    g_in = out
                    for blk in grp.residual_group:
                        t = conv(out, blk.rcab[0], slope=0.0)
                        t2 = conv(t, blk.rcab[2])
                        ca = blk.rcab[3].attention
                        mean, att = e32(B, nf), e32(B, nf)
                        w1 = ca[1].weight.detach().float().reshape(ca[1].weight.shape[0], nf).contiguous()
                        w2 = ca[3].weight.detach().float().reshape(nf, ca[3].weight.shape[1]).contiguous()
                        b1, b2 = ca[1].bias.detach().float().contiguous(), ca[3].bias.detach().float().contiguous()
                        self.keep += [w1, w2, b1, b2]
                        nxt = e16(*out.shape)
                        steps.append(lambda a=t2, m=mean: ops.channel_mean(a, m))
                        steps.append(lambda m=mean, a=att, p=(w1, b1, w2, b2): ops.ca_mlp(m, p[0], p[1], p[2], p[3], a))
                        steps.append(lambda a=t2, at=att, idn=out, o=nxt, s=float(blk.res_scale):
                                     ops.ca_scale_add(a, at, idn, o, s))
                        out = nxt
                    out = conv(out, grp.conv, res=g_in, res_scale=1.0, res_mul=1.0)
            res = conv(out, net.conv_after_body, res=x, res_scale=1.0, res_mul=1.0)
            up = res
            for i in range(0, len(net.upsample), 2):
                c = net.upsample[i]
                up = conv(up, c, ps_r=int(round(math.sqrt(c.out_channels // c.in_channels))))
            y = conv(up, net.conv_last, cout_pad=16, out_fp32=True)
            self.out = e32(B, net.num_out_ch, H * r_total, W * r_total)
            steps.append(lambda: ops.sr_output(y, self.out, 1.0 / float(net.img_range), eng.mean_dev, None, 1))
        self.graph = None

    def _build_rrdb(self, eng, B, H, W, conv, e16, e32):
        """rrdbnet_arch.RRDBNet.forward (:105-123).  Every ResidualDenseBlock works in one NHWC buffer of
        num_feat + 4 * grow channels: its input occupies the first num_feat channels, conv_k writes its grow channels
        behind them (torch.cat is a channel offset), conv5 merges x5 * 0.2 + x into the next block's buffer."""
        net, steps = eng.net, self.steps
        nf, g = net.num_feat, net.num_grow_ch
        if nf % 32 or g % 16:
            raise NotImplementedError('RRDBNet: num_feat % 32 == 0 and num_grow_ch % 16 == 0 are supported')
        s = 2 if net.scale == 2 else (4 if net.scale == 1 else 1)
        if H % s or W % s:
            raise ValueError(f'RRDBNet(scale={net.scale}) needs input extents divisible by {s}')
        h, w = H // s, W // s
        cin = net.num_in_ch * s * s
        cin_pad = -(-cin // 16) * 16
        x0 = e16(B, h, w, cin_pad)
        steps.append(lambda: ops.nchw_to_nhwc_pad(self.x_in, x0, None, 1.0, s))
        ctot = nf + 4 * g
        dense = [e16(B, h, w, ctot) for _ in range(4)]      # ring: an RRDB keeps its first buffer until its final merge

        def dconv(src, cin_used, mod, dst, c_off, **kw):
            """conv over the first cin_used channels of dense buffer `src` -> channels [c_off, c_off+cout) of `dst`."""
            wt, bias = eng.packed(mod, None, None, 0)
            cout = wt.shape[0]
            v = ops.View(src.data_ptr(), cin_used, w, h, B, src.shape[3], w * src.shape[3], h * w * src.shape[3])
            oc = dst.shape[3]
            steps.append(ops.ConvOp([v], wt, cin_used, cout, ops.taps_3x3(), (w, h, B), dst, (oc, w * oc, h * w * oc),
                                    bias=bias, out_c_off=c_off, **kw))

        feat = e16(B, h, w, nf)
        conv(x0, net.conv_first, feat, cin_pad=cin_pad)
        # feat is kept for the trunk residual; the first dense buffer starts as a copy of it (feat * 0 + feat, strided)
        steps.append(lambda: ops.ca_scale_add(feat, None, feat, dense[0], 0.0, nf, ctot))
        k = 0
        for rrdb in net.body:
            first = dense[k % len(dense)]
            cur = first
            for j, rdb in enumerate((rrdb.rdb1, rrdb.rdb2, rrdb.rdb3)):
                nxt = dense[(k + 1) % len(dense)]
                for i, c in enumerate((rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4)):
                    dconv(cur, nf + i * g, c, cur, nf + i * g, act_slope=0.2)
                res_kw = dict(res=cur, res_mode=1, res_strides=(ctot, w * ctot, h * w * ctot), res_wh=(w, h),
                              res_scale=0.2, res_mul=1.0)
                if j < 2:
                    dconv(cur, ctot, rdb.conv5, nxt, 0, **res_kw)                      # x5 * 0.2 + x -> next block's input
                else:
                    t = e16(B, h, w, nf)
                    dconv(cur, ctot, rdb.conv5, t, 0, **res_kw)
                    steps.append(lambda a=t, idn=first, o=nxt: ops.ca_scale_add(a, None, idn, o, 0.2, ctot, ctot))  # RRDB: out*0.2 + x
                cur = nxt
                k += 1
        body_out = cur                                                                   # first nf channels
        trunk = e16(B, h, w, nf)
        wt, bias = eng.packed(net.conv_body, None, None, 0)
        v = ops.View(body_out.data_ptr(), nf, w, h, B, ctot, w * ctot, h * w * ctot)
        steps.append(ops.ConvOp([v], wt, nf, nf, ops.taps_3x3(), (w, h, B), trunk, (nf, w * nf, h * w * nf), bias=bias,
                                res=feat, res_mode=1, res_strides=(nf, w * nf, h * w * nf), res_wh=(w, h), res_scale=1.0,
                                res_mul=1.0))                                            # feat + conv_body(body(feat))
        up = trunk
        for c in (net.conv_up1, net.conv_up2):
            bb, hh, ww, _ = up.shape
            big = e16(bb, 2 * hh, 2 * ww, nf)
            steps.append(lambda a=up, o=big: ops.nearest_up2(a, o))
            up = conv(big, c, slope=0.2)
        hr = conv(up, net.conv_hr, slope=0.2)
        y = conv(hr, net.conv_last, cout_pad=16, out_fp32=True)
        self.out = e32(B, net.num_out_ch, y.shape[1], y.shape[2])
        steps.append(lambda: ops.sr_output(y, self.out, 1.0, None, None, 1))

    def launch(self):
        for st in self.steps:
            st()


class SrEngine:
    def __init__(self, net):
        self.dev = next(net.parameters()).device
        if self.dev.type != 'cuda':
            raise RuntimeError('image_restoration_b200 needs the module on a CUDA B200 (no CPU path)')
        from . import _lib
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().b200ir_device_check(), 'device check')
        if (net.num_in_ch > 16 and not isinstance(net, RRDBNet)) or net.num_out_ch > 16 or net.num_feat % 16:
            raise NotImplementedError('num_in_ch / num_out_ch <= 16 and num_feat % 16 == 0 are supported')
        self.net = net
        self._sig = self._signature()
        self._packed = {}
        self.plans = {}
        self.use_graphs = True
        self.mean_dev = net.mean.reshape(-1).to(self.dev, F32).contiguous() if hasattr(net, 'mean') else None

    def _signature(self):
        ps = list(self.net.parameters())
        return (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))

    def stale(self):
        return self._sig != self._signature()

    def packed(self, mod, cin_pad, cout_pad, ps_r):
        key = (id(mod), cin_pad, cout_pad, ps_r)
        if key not in self._packed:
            self._packed[key] = _pack(mod, cin_pad, cout_pad, ps_r)
        return self._packed[key]

    @torch.no_grad()
    def forward(self, x):
        net = self.net
        if x.dim() != 4 or x.shape[1] != net.num_in_ch:
            raise ValueError(f'expected input (B,{net.num_in_ch},H,W), got {tuple(x.shape)}')
        B, _, H, W = x.shape
        with torch.cuda.device(self.dev):
            key = (B, H, W)
            if key in self.plans:
                self.plans[key] = self.plans.pop(key)
            else:
                while len(self.plans) >= 4:             # keep the four most recently used input shapes
                    self.plans.pop(next(iter(self.plans)))
                self.plans[key] = _SrPlan(self, B, H, W)
            plan = self.plans[key]
            plan.x_in.copy_(x)
            if not self.use_graphs:
                plan.launch()
            else:
                if plan.graph is None:
                    plan.launch()
                    torch.cuda.synchronize()
                    plan.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(plan.graph):
                        plan.launch()
                plan.graph.replay()
            out = plan.out.clone()
        return out.to(x.dtype) if x.dtype != F32 else out


def register_sr_archs(registry=None, suffix='_B200', override=False):
    """Adds MSRResNet / EDSR / RCAN to a basicsr-style registry under `<name><suffix>` (YAML: `type: MSRResNet_B200`),
    or replaces the reference entries when `override` (Registry asserts on duplicates, registry.py:38-41)."""
    registry = registry if registry is not None else ARCH_REGISTRY
    out = {}
    for cls in (MSRResNet, EDSR, RCAN, RRDBNet):
        if override:
            registry._obj_map[cls.__name__] = cls
            out[cls.__name__] = cls
        else:
            name = cls.__name__ + suffix
            if name not in registry:
                registry.register(type(name, (cls,), {'__doc__': cls.__doc__}))
            out[name] = registry.get(name)
    return out


register_sr_archs(ARCH_REGISTRY)
if not USING_BASICSR_REGISTRY:
    for _cls in (MSRResNet, EDSR, RCAN, RRDBNet):
        if _cls.__name__ not in ARCH_REGISTRY:
            ARCH_REGISTRY.register(_cls)
