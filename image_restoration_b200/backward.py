"""Backward pass of the optional training step (BASELINE config 5): torch.autograd.Function wrappers in
the way the reference wraps its own native ops (FusedLeakyReLUFunction, basicsr/ops/fused_act/fused_act.py:66-79: the
Python wrapper owns save-for-backward, the native side does the arithmetic), and the two networks assembled from them.

    ConvLayer3x3Function   EqualConv2d 3x3 + FusedLeakyReLU / + bias only        stylegan2_ocr_arch.py:658-705, SFT heads
    ResBlockFunction       conv1, FIR + stride-2 conv2, FIR + 1x1 skip            stylegan2_ocr_arch.py:708-734
    ResUpBlockFunction     conv1, bilinear x2 + conv2, 1x1 skip + bilinear x2     gfpganv1_ocr_arch.py:205-225
    EqualLinearFunction    EqualLinear (+ fused_lrelu)                            stylegan2_ocr_arch.py:165-175
    FirstConvFunction      conv_body_first (1x1 over the fp32 NCHW image)
    ToRGBHeadFunction      toRGB[i] heads of the image pyramid loss               gfpganv1_ocr_arch.py:308-311, 377-378
    MinibatchStddevFunction                                                       stylegan2_arch.py:791-801
    unet_forward           trainable part of GFPGANv1OCR.forward                  gfpganv1_ocr_arch.py:352-378
    disc_forward           StyleGAN2Discriminator.forward (network_d)             stylegan2_arch.py:788-805

For one ConvLayer:
    forward   y  = lrelu(conv2d(x, W / sqrt(9 cin), padding=1) + b, 0.2) * sqrt 2    b200ir_conv_igemm (fused epilogue)
    backward  dz = dy * sqrt 2 * (y > 0 ? 1 : 0.2),  db = sum dz                      b200ir_lrelu_bias_bwd
              dW = x (*) dz / sqrt(9 cin)                                              b200ir_conv_wgrad (tcgen05, MN-major)
              dx = conv2d(dz, flip(W)^T, padding=1)                                    b200ir_conv_igemm (adjoint weights)

Activations and their gradients are NHWC fp16 on the device (use a loss scale when the loss gradient is small: see
tests/test_train_step_gpu.py); parameters and their gradients are fp32 in the reference's layouts.  torch packs weights,
allocates buffers and runs the autograd graph; no CPU path: every arithmetic call lands in libb200ir.so.
"""
import math

import torch

from . import _lib, ops


def _need_cuda(t, who):
    _lib.require_cuda(t, f'backward.{who}')


# A forward pass can be tagged (`with graph_tag('net_d'):`); a later backward pass through those nodes can then be told to
# leave the parameter gradients out (`with skip_param_grads('net_d'):`).  optimize_parameters runs net_d on the generator's
# output twice with identical weights — once for l_g_gan (input gradient only) and once for l_d (parameter gradients only,
# gfpgan_model.py:549-552, :676-681); the trainer runs that forward ONCE and walks its graph twice (train.GFPGANTrainer).
_TAG = [None]
_SKIP_PARAM_GRADS = set()


class graph_tag:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        self.prev, _TAG[0] = _TAG[0], self.tag

    def __exit__(self, *exc):
        _TAG[0] = self.prev
        return False


class skip_param_grads:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        _SKIP_PARAM_GRADS.add(self.tag)

    def __exit__(self, *exc):
        _SKIP_PARAM_GRADS.discard(self.tag)
        return False


def _need(ctx, params):
    """ctx.needs_input_grad with the parameter positions cleared when this node's tag is in the skip set."""
    need = ctx.needs_input_grad
    if getattr(ctx, 'tag', None) in _SKIP_PARAM_GRADS:
        need = tuple(False if i in params else n for i, n in enumerate(need))
    return need


def pack_equal_conv3x3(weight):
    """EqualConv2d weight fp32 [cout, cin, 3, 3] -> (packed fp16 [cout, 9*cin] with the equalised-lr scale folded in,
    scale).  stylegan2_ocr_arch.py:629-648."""
    cout, cin, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    scale = 1.0 / math.sqrt(cin * 9)
    return (weight.detach() * scale).permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(torch.float16).contiguous(), scale


class ConvLayer3x3Function(torch.autograd.Function):
    """y = FusedLeakyReLU(EqualConv2d(x)) (activate=True; also EqualConv2d(bias=True) + ScaledLeakyReLU of the SFT heads,
    gfpganv1_ocr_arch.py:322-339) or y = EqualConv2d(x) + bias (activate=False) on NHWC fp16 activations; gradients for x,
    weight and bias."""

    @staticmethod
    def forward(ctx, x, weight, bias, activate=True):
        _need_cuda(x, 'ConvLayer3x3Function')
        b, h, w, cin = x.shape
        cout = weight.shape[0]
        wp, scale = pack_equal_conv3x3(weight)
        y = torch.empty(b, h, w, cout, device=x.device, dtype=torch.float16)
        ops.conv_same(x, wp, y, 3, bias=bias.detach().float().contiguous(), act=activate)()
        ctx.save_for_backward(x, y if activate else None, wp)
        ctx.scale = scale
        ctx.tag = _TAG[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, wp = ctx.saved_tensors
        need = _need(ctx, (1, 2))
        b, h, w, cin = x.shape
        cout = wp.shape[0]
        dy = dy.contiguous()
        if y is not None:
            dz, dbias = ops.lrelu_bias_bwd(dy, y, want_bias=need[2])
        else:
            dz, dbias = dy, (ops.lrelu_bias_bwd(dy, None, scale=1.0)[1] if need[2] else None)
        dx = dweight = None
        if need[1]:
            dw = ops.conv_wgrad(x, dz)                                           # [cout, 9, cin] fp32
            dweight = dw.view(cout, 3, 3, cin).permute(0, 3, 1, 2) * ctx.scale   # reference layout [cout, cin, 3, 3]
        if need[0]:
            dx = torch.empty_like(x)
            ops.conv_dgrad(dz, ops.conv_dgrad_weight(wp, cin), dx)()
        return dx, dweight, dbias, None


def conv_layer3x3(x, weight, bias, activate=True):
    return ConvLayer3x3Function.apply(x, weight, bias, activate)


class EqualLinearFunction(torch.autograd.Function):
    """EqualLinear without activation (stylegan2_ocr_arch.py:165-175; final_linear of GFPGANv1OCR, 12288 -> 3072):
    y = x @ (W * scale)^T + bias * lr_mul on fp16 rows x [B, in].  Forward and input gradient are 1x1 convs over B
    'pixels'; the weight gradient dW = dy^T x is the wgrad kernel with the batch laid out along the pixel axis."""

    @staticmethod
    def forward(ctx, x, weight, bias, lr_mul=1.0, activate=False):
        _need_cuda(x, 'EqualLinearFunction')
        b, cin = x.shape
        cout = weight.shape[0]
        scale = lr_mul / math.sqrt(cin)
        wp = (weight.detach() * scale).to(torch.float16).contiguous()
        y = torch.empty(b, cout, device=x.device, dtype=torch.float16)
        # activate: EqualLinear(activation='fused_lrelu') = fused_leaky_relu(out, bias * lr_mul) (stylegan2_arch.py:165-175)
        ops.conv_same(x.view(1, 1, b, cin), wp, y.view(1, 1, b, cout), 1, bias=(bias.detach().float() * lr_mul).contiguous(),
                      act=activate)()
        ctx.save_for_backward(x, wp, y if activate else None)
        ctx.consts = (scale, lr_mul)
        ctx.tag = _TAG[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp, y = ctx.saved_tensors
        need = _need(ctx, (1, 2))
        scale, lr_mul = ctx.consts
        b, cin = x.shape
        cout = wp.shape[0]
        dy = dy.contiguous()
        dx = dweight = dbias = None
        if y is not None:
            dy, dbias = ops.lrelu_bias_bwd(dy, y, want_bias=need[2])
            dbias = dbias * lr_mul if dbias is not None else None
        elif need[2]:
            dbias = ops.lrelu_bias_bwd(dy, None, scale=1.0)[1] * lr_mul
        if need[1]:
            dweight = ops.conv1x1_wgrad(x.view(1, 1, b, cin), dy.view(1, 1, b, cout)) * scale
        if need[0]:
            dx = torch.empty_like(x)
            ops.conv_same(dy.view(1, 1, b, cout), wp.t().contiguous(), dx.view(1, 1, b, cin), 1)()
        return dx, dweight, dbias, None, None


def equal_linear(x, weight, bias, lr_mul=1.0, activate=False):
    return EqualLinearFunction.apply(x, weight, bias, lr_mul, activate)


def pack_equal_conv(weight):
    """EqualConv2d weight fp32 [cout, cin, k, k] -> packed fp16 [cout, k*k*cin] (tap-major) with the equalised-lr scale."""
    cout, cin, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    return (weight.detach() * scale).permute(0, 2, 3, 1).reshape(cout, k * k * cin).to(torch.float16).contiguous(), scale


class ResBlockFunction(torch.autograd.Function):
    """ResBlock of the U-Net encoder (stylegan2_ocr_arch.py:708-734) on NHWC fp16 activations, forward and backward:

        t1  = FusedLeakyReLU(conv3x3(x, W1) + b1)                      conv1
        y2  = FusedLeakyReLU(conv3x3_s2(FIR_pad22(t1), W2) + b2)       conv2 (downsample=True)
        out = (y2 + conv1x1(FIR_down2(x), Ws)) / sqrt 2                skip (downsample=True, no bias, no activation)

    Backward = the adjoint chain on the same kernels: lrelu_bias_bwd, wgrad (plain / per-phase / 1x1), the merged
    transposed conv (dgrad of the stride-2 conv) + fir_pad11, conv_dgrad, 1x1 dgrad + fir_down2_adjoint (which also adds
    the two branches of dx)."""

    debug_saved = None

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ws):
        _need_cuda(x, 'ResBlockFunction')
        out, saved, scales = resblock_forward(x, w1, b1, w2, b2, ws)
        ctx.save_for_backward(*saved)
        ctx.scales = scales
        ctx.tag = _TAG[0]
        if ResBlockFunction.debug_saved is not None:      # tests read the leaky-ReLU branches the kernels took
            ResBlockFunction.debug_saved.update(t1=saved[1], y2=saved[3])
        return out

    @staticmethod
    def backward(ctx, dout):
        return resblock_backward(ctx.saved_tensors, ctx.scales, dout, _need(ctx, (1, 2, 3, 4, 5)))[:6]


def resblock_forward(x, w1, b1, w2, b2, ws):
    """Forward launches of one ResBlock (see ResBlockFunction).  Returns (out, saved, scales) with
    saved = (x, t1, p, y2, sk_in, w1p, w2, wsp): what the backward pass — and the tangent pass of the R1 penalty (r1.py) — needs."""
    b, h, w, cin = x.shape
    cout = w2.shape[0]
    oh, ow = h // 2, w // 2
    dev = x.device
    w1p, s1 = pack_equal_conv(w1)
    w2p, s2 = pack_equal_conv(w2)
    wsp, ss = pack_equal_conv(ws)
    e16 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float16)   # noqa: E731
    t1 = e16(b, h, w, cin)
    ops.conv_same(x, w1p, t1, 3, bias=b1.detach().float().contiguous(), act=True)()
    p = smoothed_buffer(t1)
    y2 = e16(b, oh, ow, cout)
    ops.conv3x3_s2(p, h, w, w2p, y2, bias=b2.detach().float().contiguous(), act=True)()
    sk_in = e16(b, oh, ow, cin)
    ops.fir_down2(x, sk_in)
    out = e16(b, oh, ow, cout)
    ops.conv_same(sk_in, wsp, out, 1, res=y2, res_mode=1, res_strides=(cout, ow * cout, oh * ow * cout), res_wh=(ow, oh),
                  res_scale=ops.INV_SQRT2)()
    return out, (x, t1, p, y2, sk_in, w1p, w2.detach(), wsp), (s1, s2, ss)


def smoothed_buffer(t1):
    """fir_pad22(t1) into a fresh [B, h+2, w+2, C] buffer (the input of the stride-2 conv over phase views).  fir_pad22 writes
    rows 0..h / columns 0..w; the spare row h+1 and column w+1 only need to be FINITE: the weight gradient's contraction runs
    over pixels, and the TMA box of a ragged tile pairs them with zero-filled dy (0 x NaN)."""
    b, h, w, c = t1.shape
    p = torch.empty(b, h + 2, w + 2, c, device=t1.device, dtype=torch.float16)
    p[:, h + 1].zero_()
    p[:, :, w + 1].zero_()
    ops.fir_pad22(t1, p)
    return p


def resblock_backward(saved, scales, dout, need):
    """Backward launches of one ResBlock; need = (x, w1, b1, w2, b2, ws) flags.  Returns (dx, dw1, db1, dw2, db2, dws, dz1, dz2):
    the last two are the gradients at the pre-activations of conv1 / conv2, which the R1 penalty pairs with tangents."""
    x, t1, p, y2, sk_in, w1p, w2, wsp = saved
    s1, s2, ss = scales
    b, h, w, cin = x.shape
    oh, ow, cout = y2.shape[1:]
    dev = x.device
    dout = dout.contiguous()
    # conv2: the 1 / sqrt 2 of the merge and the sqrt 2 of FusedLeakyReLU cancel
    dz2, db2 = ops.lrelu_bias_bwd(dout, y2, scale=1.0, want_bias=need[4])
    dw2 = None
    if need[3]:
        dw2 = ops.conv3x3_s2_wgrad(p, h, w, dz2).view(cout, 3, 3, cin).permute(0, 3, 1, 2) * s2
    # dgrad of the stride-2 conv = conv_transpose2d(dz2, W2, stride 2) -> (h+1) x (w+1), then the FIR adjoint
    raw = torch.empty(b, h + 2, w + 2, cin, device=dev, dtype=torch.float16)
    ops.convt_s2_merged(dz2, ops.convt_merged_weight(w2.detach().permute(1, 0, 2, 3), s2), raw, None)()
    dt1 = torch.empty_like(t1)
    ops.fir_pad11(raw, dt1)
    # conv1
    dz1, db1 = ops.lrelu_bias_bwd(dt1, t1, dz=dt1, want_bias=need[2])
    dw1 = ops.conv_wgrad(x, dz1).view(cin, 3, 3, cin).permute(0, 3, 1, 2) * s1 if need[1] else None
    # skip: d(conv1x1 input) with the 1 / sqrt 2 folded into the weights
    dws = None
    if need[5]:
        dws = (ops.conv1x1_wgrad(sk_in, dout) * (ss * ops.INV_SQRT2)).reshape(cout, cin, 1, 1)
    dx = None
    if need[0]:
        dx = torch.empty_like(x)
        ops.conv_dgrad(dz1, ops.conv_dgrad_weight(w1p, cin), dx)()
        dsk = torch.empty_like(sk_in)
        wst = (wsp.float() * ops.INV_SQRT2).t().contiguous().to(torch.float16)          # [cin, cout]
        ops.conv_same(dout, wst, dsk, 1)()
        ops.fir_down2_adjoint(dsk, dx, add=dx)
    return dx, dw1, db1, dw2, db2, dws, dz1, dz2


def res_block(x, w1, b1, w2, b2, ws):
    return ResBlockFunction.apply(x, w1, b1, w2, b2, ws)


class ResUpBlockFunction(torch.autograd.Function):
    """ResUpBlock of the U-Net decoder (gfpganv1_ocr_arch.py:205-225) on NHWC fp16 activations, forward and backward:

        t1  = FusedLeakyReLU(conv3x3(x, W1) + b1)                                conv1
        y2  = FusedLeakyReLU(conv3x3(bilinear_up2(t1), W2) + b2)                 conv2 (ConvUpLayer)
        out = (y2 + bilinear_up2(conv1x1(x, Ws))) / sqrt 2                       skip (the 1x1 conv commutes with bilinear x2)

    y2 / sqrt 2 is produced directly (leaky ReLU without the sqrt 2 gain) and kept for the backward pass; the 1 / sqrt 2
    of the skip is folded into Ws."""

    debug_saved = None

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ws):
        _need_cuda(x, 'ResUpBlockFunction')
        b, h, w, cin = x.shape
        cout = w2.shape[0]
        dev = x.device
        w1p, s1 = pack_equal_conv(w1)
        w2p, s2 = pack_equal_conv(w2)
        wsp, ss = pack_equal_conv(ws * ops.INV_SQRT2)
        e16 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float16)   # noqa: E731
        t1 = e16(b, h, w, cin)
        ops.conv_same(x, w1p, t1, 3, bias=b1.detach().float().contiguous(), act=True)()
        u = e16(b, 2 * h, 2 * w, cin)
        ops.bilinear_up2(t1, u)
        y2 = e16(b, 2 * h, 2 * w, cout)
        ops.conv_same(u, w2p, y2, 3, bias=b2.detach().float().contiguous(), act_slope=0.2)()
        sl = e16(b, h, w, cout)
        ops.conv_same(x, wsp, sl, 1)()
        su = e16(b, 2 * h, 2 * w, cout)
        ops.bilinear_up2(sl, su)
        out = e16(b, 2 * h, 2 * w, cout)
        ops.add(y2, su, out)
        ctx.save_for_backward(x, t1, u, y2, w1p, w2p, wsp)
        ctx.scales = (s1, s2, ss * ops.INV_SQRT2)
        if ResUpBlockFunction.debug_saved is not None:
            ResUpBlockFunction.debug_saved.update(t1=t1, y2=y2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, t1, u, y2, w1p, w2p, wsp = ctx.saved_tensors
        s1, s2, ss = ctx.scales
        b, h, w, cin = x.shape
        cout = y2.shape[3]
        need = ctx.needs_input_grad
        dout = dout.contiguous()
        dz2, db2 = ops.lrelu_bias_bwd(dout, y2, scale=1.0, want_bias=need[4])
        dw2 = ops.conv_wgrad(u, dz2).view(cout, 3, 3, cin).permute(0, 3, 1, 2) * s2 if need[3] else None
        du = torch.empty_like(u)
        ops.conv_dgrad(dz2, ops.conv_dgrad_weight(w2p, cin), du)()
        dt1 = torch.empty_like(t1)
        ops.bilinear_up2_adjoint(du, dt1)
        dz1, db1 = ops.lrelu_bias_bwd(dt1, t1, dz=dt1, want_bias=need[2])
        dw1 = ops.conv_wgrad(x, dz1).view(cin, 3, 3, cin).permute(0, 3, 1, 2) * s1 if need[1] else None
        dsl = torch.empty(b, h, w, cout, device=x.device, dtype=torch.float16)
        ops.bilinear_up2_adjoint(dout, dsl)
        dws = (ops.conv1x1_wgrad(x, dsl) * ss).reshape(cout, cin, 1, 1) if need[5] else None
        dx = None
        if need[0]:
            dx1 = torch.empty_like(x)
            ops.conv_dgrad(dz1, ops.conv_dgrad_weight(w1p, cin), dx1)()
            dx = torch.empty_like(x)
            ops.conv_same(dsl, wsp.t().contiguous(), dx, 1, res=dx1, res_mode=1, res_strides=(cin, w * cin, h * w * cin),
                          res_wh=(w, h), res_scale=1.0)()
        return dx, dw1, db1, dw2, db2, dws


def res_up_block(x, w1, b1, w2, b2, ws):
    return ResUpBlockFunction.apply(x, w1, b1, w2, b2, ws)


class FirstConvFunction(torch.autograd.Function):
    """conv_body_first = ConvLayer(3, C, 1, bias=True, activate=True) over the fp32 NCHW input image.  The image is data for
    net_g (no input gradient); the discriminator's conv_body.0 returns d/d(image) when the image requires grad — that is how
    l_g_gan reaches net_g (gfpgan_model.py:549-552)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, 'FirstConvFunction')
        b, _, h, w = x.shape
        cout = weight.shape[0]
        scale = 1.0 / math.sqrt(3.0)
        y = torch.empty(b, h, w, cout, device=x.device, dtype=torch.float16)
        ws = (weight.detach().view(cout, 3) * scale).contiguous()
        xc = x.detach().contiguous()
        ops.first_conv(xc, ws, bias.detach().float().contiguous(), y)
        ctx.save_for_backward(xc, y, ws)
        ctx.scale = scale
        ctx.tag = _TAG[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, ws = ctx.saved_tensors
        need = _need(ctx, (1, 2))
        dz, dbias = ops.lrelu_bias_bwd(dy.contiguous(), y, want_bias=need[2])
        dw = (ops.first_conv_wgrad(x, dz) * ctx.scale).view(-1, 3, 1, 1) if need[1] else None
        dx = None
        if need[0]:
            dx = torch.empty_like(x)
            ops.first_conv_dgrad(dz, ws, dx)
        return dx, dw, dbias


class AddFunction(torch.autograd.Function):
    """feat + unet_skips[i] (gfpganv1_ocr_arch.py:368) on the b200ir_add kernel; the gradient goes to both inputs."""

    @staticmethod
    def forward(ctx, a, b):
        _need_cuda(a, 'AddFunction')
        out = torch.empty_like(a)
        ops.add(a, b, out)
        return out

    @staticmethod
    def backward(ctx, d):
        return d, d


class ToRGBHeadFunction(torch.autograd.Function):
    """toRGB[i] of the U-Net (EqualConv2d(C, 3, 1) with bias, gfpganv1_ocr_arch.py:308-311, 377-378), the heads of the image
    pyramid loss: the three output channels are padded to 16 so that the 1x1 conv, its input gradient and its weight
    gradient run on the GEMM kernels.  Returns NHWC fp16 [B,h,w,16]; channels 3.. are zero and carry no gradient."""

    PAD = 16

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, 'ToRGBHeadFunction')
        b, h, w, cin = x.shape
        pad = ToRGBHeadFunction.PAD
        scale = 1.0 / math.sqrt(cin)
        wp = torch.zeros(pad, cin, device=x.device, dtype=torch.float16)
        wp[:3] = (weight.detach().view(3, cin) * scale).to(torch.float16)
        bp = torch.zeros(pad, device=x.device, dtype=torch.float32)
        bp[:3] = bias.detach().float()
        y = torch.empty(b, h, w, pad, device=x.device, dtype=torch.float16)
        ops.conv_same(x, wp, y, 1, bias=bp)()
        ctx.save_for_backward(x, wp)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp = ctx.saved_tensors
        cin = x.shape[3]
        dy = dy.contiguous()
        dbias = ops.lrelu_bias_bwd(dy, None, scale=1.0)[1][:3]
        dweight = (ops.conv1x1_wgrad(x, dy)[:3] * ctx.scale).reshape(3, cin, 1, 1)
        dx = torch.empty_like(x)
        ops.conv_same(dy, wp.t().contiguous(), dx, 1)()
        return dx, dweight, dbias


def unet_forward(sd, x, different_w=True, num_style_feat=256, return_rgb=False):
    """The trainable part of GFPGANv1OCR.forward (gfpganv1_ocr_arch.py:352-378; everything the optimiser touches when
    fix_decoder=True): U-Net encoder -> style code, U-Net decoder -> SFT conditions, on the B200 kernels with autograd
    through the Function wrappers above.  `sd`: tensors under the reference's state_dict names (fp32 CUDA parameters);
    x fp32 NCHW [B,3,H,W].  Returns (style_code fp16 [B, num_latent, num_style_feat] or [B, n], conditions: list of NHWC fp16
    tensors scale0, shift0, scale1, ...); with return_rgb also out_rgbs, the toRGB heads of the image pyramid loss
    (gfpgan_model.py:531-536), NHWC fp16 [B,h,w,16] with the image in channels 0..2."""
    feat = FirstConvFunction.apply(x.contiguous(), sd['conv_body_first.0.weight'], sd['conv_body_first.1.bias'])
    levels = 0
    while f'conv_body_down.{levels}.conv1.0.weight' in sd:
        levels += 1
    skips = []
    for i in range(levels):
        pre = f'conv_body_down.{i}'
        feat = res_block(feat, sd[f'{pre}.conv1.0.weight'], sd[f'{pre}.conv1.1.bias'], sd[f'{pre}.conv2.1.weight'],
                         sd[f'{pre}.conv2.2.bias'], sd[f'{pre}.skip.1.weight'])
        skips.insert(0, feat)
    feat = conv_layer3x3(feat, sd['final_conv.0.weight'], sd['final_conv.1.bias'])
    b, h, w, c = feat.shape
    # final_linear reads the NCHW flattening (gfpganv1_ocr_arch.py:361): permute its columns to the NHWC order instead
    wl = sd['final_linear.weight']
    wl = wl.view(wl.shape[0], c, h, w).permute(0, 2, 3, 1).reshape(wl.shape[0], -1)
    style_code = equal_linear(feat.reshape(b, -1), wl, sd['final_linear.bias'])
    if different_w:
        style_code = style_code.view(b, -1, num_style_feat)
    conditions, out_rgbs = [], []
    for i in range(levels):
        pre = f'conv_body_up.{i}'
        feat = AddFunction.apply(feat, skips[i])
        feat = res_up_block(feat, sd[f'{pre}.conv1.0.weight'], sd[f'{pre}.conv1.1.bias'], sd[f'{pre}.conv2.weight'],
                            sd[f'{pre}.conv2.activation.bias'], sd[f'{pre}.skip.weight'])
        for head in ('condition_scale', 'condition_shift'):
            hid = conv_layer3x3(feat, sd[f'{head}.{i}.0.weight'], sd[f'{head}.{i}.0.bias'])
            conditions.append(conv_layer3x3(hid, sd[f'{head}.{i}.2.weight'], sd[f'{head}.{i}.2.bias'], False))
        if return_rgb:
            out_rgbs.append(ToRGBHeadFunction.apply(feat, sd[f'toRGB.{i}.weight'], sd[f'toRGB.{i}.bias']))
    if return_rgb:
        return style_code, conditions, out_rgbs
    return style_code, conditions



class MinibatchStddevFunction(torch.autograd.Function):
    """Minibatch standard deviation of the discriminator (stylegan2_arch.py:791-801): x NHWC fp16 [B,h,w,C] -> [B,h,w,c_pad]
    = x, the group statistic in channel C, zeros up to c_pad (a multiple of 16, the next conv's input block)."""

    @staticmethod
    def forward(ctx, x, group):
        _need_cuda(x, 'MinibatchStddevFunction')
        b, h, w, c = x.shape
        c_pad = (c + 1 + 15) // 16 * 16
        out = torch.zeros(b, h, w, c_pad, device=x.device, dtype=torch.float16)
        s_buf = torch.empty(b // group, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().b200ir_minibatch_stddev(ops._ptr(x), ops._ptr(s_buf), ops._ptr(out), b, h * w, c, c_pad, group,
                                                      ops._stream()), 'minibatch_stddev')
        ctx.save_for_backward(x)
        ctx.group = group
        return out

    @staticmethod
    def backward(ctx, dcat):
        (x,) = ctx.saved_tensors
        b, h, w, c = x.shape
        g = ctx.group
        dcat = dcat.contiguous()
        ds = dcat[..., c].float().view(g, b // g, h * w).sum(dim=(0, 2)).contiguous()      # B * h * w numbers: host-side glue
        dx = torch.empty_like(x)
        _lib.check(_lib.lib().b200ir_minibatch_stddev_bwd(ops._ptr(x), ops._ptr(dcat), ops._ptr(ds), ops._ptr(dx), b, h * w, c,
                                                          dcat.shape[3], g, ops._stream()), 'minibatch_stddev_bwd')
        return dx, None


def disc_forward(sd, x, stddev_group=4):
    """StyleGAN2Discriminator.forward (stylegan2_arch.py:788-805; network_d of the training YAMLs) with autograd through
    the Function wrappers: `sd` = fp32 CUDA parameters under the reference's state_dict names, x fp32 NCHW [B,3,H,W] ->
    scores fp16 [B, 1]."""
    feat = FirstConvFunction.apply(x.contiguous(), sd['conv_body.0.0.weight'], sd['conv_body.0.1.bias'])
    i = 1
    while f'conv_body.{i}.conv1.0.weight' in sd:
        pre = f'conv_body.{i}'
        feat = res_block(feat, sd[f'{pre}.conv1.0.weight'], sd[f'{pre}.conv1.1.bias'], sd[f'{pre}.conv2.1.weight'],
                         sd[f'{pre}.conv2.2.bias'], sd[f'{pre}.skip.1.weight'])
        i += 1
    b, h, w, c = feat.shape
    group = min(b, stddev_group)
    cat = MinibatchStddevFunction.apply(feat, group)
    c_pad = cat.shape[3]
    wf = sd['final_conv.0.weight']                                       # [c4, C + 1, 3, 3]: zero input channels up to c_pad
    wf = torch.nn.functional.pad(wf, (0, 0, 0, 0, 0, c_pad - wf.shape[1])) * math.sqrt(c_pad / (c + 1.0))   # keeps 1/sqrt((C+1)*9)
    g = conv_layer3x3(cat, wf, sd['final_conv.1.bias'])
    c4 = g.shape[3]
    w1 = sd['final_linear.0.weight']                                     # reads the NCHW flattening: permute its columns
    w1 = w1.view(w1.shape[0], c4, h * w).permute(0, 2, 1).reshape(w1.shape[0], -1)
    hid = equal_linear(g.reshape(b, -1), w1, sd['final_linear.0.bias'], 1.0, True)
    w2 = torch.nn.functional.pad(sd['final_linear.1.weight'], (0, 0, 0, 15))   # one output row, padded to the 16-channel block
    b2 = torch.nn.functional.pad(sd['final_linear.1.bias'], (0, 15))
    return equal_linear(hid, w2, b2)[:, :1]
