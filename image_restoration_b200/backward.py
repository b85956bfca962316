"""Backward pass pieces of the optional training step (BASELINE config 5) for the U-Net encoder's
`ConvLayer(cin, cout, 3, bias=True, activate=True)` (stylegan2_ocr_arch.py:658-705: EqualConv2d 3x3 stride 1 pad 1
followed by FusedLeakyReLU), as a torch.autograd.Function in the way the reference wraps its own native ops
(FusedLeakyReLUFunction, basicsr/ops/fused_act/fused_act.py:66-79: the Python wrapper owns save-for-backward, the
native side does the arithmetic).

    forward   y  = lrelu(conv2d(x, W / sqrt(9 cin), padding=1) + b, 0.2) * sqrt 2    b200ir_conv_igemm (fused epilogue)
    backward  dz = dy * sqrt 2 * (y > 0 ? 1 : 0.2),  db = sum dz                      b200ir_lrelu_bias_bwd
              dW = x (*) dz / sqrt(9 cin)                                              b200ir_conv_wgrad (tcgen05, MN-major)
              dx = conv2d(dz, flip(W)^T, padding=1)                                    b200ir_conv_igemm (adjoint weights)

Activations are NHWC fp16 on the device; parameters and their gradients are fp32 in the reference's layouts
([cout, cin, 3, 3] and [cout]).  No CPU path: every call lands in libb200ir.so.
"""
import math

import torch

from . import ops


def pack_equal_conv3x3(weight):
    """EqualConv2d weight fp32 [cout, cin, 3, 3] -> (packed fp16 [cout, 9*cin] with the equalised-lr scale folded in,
    scale).  stylegan2_ocr_arch.py:629-648."""
    cout, cin, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    scale = 1.0 / math.sqrt(cin * 9)
    return (weight.detach() * scale).permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(torch.float16).contiguous(), scale


class ConvLayer3x3Function(torch.autograd.Function):
    """y = FusedLeakyReLU(EqualConv2d(x)) on NHWC fp16 activations; gradients for x, weight and bias."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if not x.is_cuda:
            raise RuntimeError('image_restoration_b200.backward.ConvLayer3x3Function needs CUDA tensors (no CPU path)')
        b, h, w, cin = x.shape
        cout = weight.shape[0]
        wp, scale = pack_equal_conv3x3(weight)
        y = torch.empty(b, h, w, cout, device=x.device, dtype=torch.float16)
        ops.conv_same(x, wp, y, 3, bias=bias.detach().float().contiguous(), act=True)()
        ctx.save_for_backward(x, y, wp)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, wp = ctx.saved_tensors
        b, h, w, cin = x.shape
        cout = y.shape[3]
        dz, dbias = ops.lrelu_bias_bwd(dy.contiguous(), y, want_bias=ctx.needs_input_grad[2])
        dx = dweight = None
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(x, dz)                                           # [cout, 9, cin] fp32
            dweight = dw.view(cout, 3, 3, cin).permute(0, 3, 1, 2) * ctx.scale   # reference layout [cout, cin, 3, 3]
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.conv_dgrad(dz, ops.conv_dgrad_weight(wp, cin), dx)()
        return dx, dweight, dbias


def conv_layer3x3(x, weight, bias):
    return ConvLayer3x3Function.apply(x, weight, bias)
