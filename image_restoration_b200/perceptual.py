"""Perceptual + style loss of the training step on the B200 kernels:

    PerceptualLoss            Car_Plate-Restoration/basicsr/losses/losses.py:250-356   (criterion 'l1', Gram-matrix style term)
    VGGFeatureExtractor       Car_Plate-Restoration/basicsr/archs/vgg_arch.py:56-160   (torchvision vgg19.features, taps before ReLU)
    call site                 basicsr/models/gfpgan_model.py:538-545  (cri_perceptual(self.output, self.gt))

The VGG19 convs run on b200ir_conv_igemm (nn.Conv2d 3x3 + bias [+ ReLU] in the epilogue), ReLU + MaxPool on
b200ir_maxpool2_relu, the feature L1 on b200ir_l1_loss_f16, the Gram matrices on the weight-gradient GEMM
(b200ir_gram_batched: sum over pixels of f (x) f for every image in one launch) and the backward pass — input gradients only, the
VGG is frozen — on the same conv kernel with adjoint weights.  The input normalisation ((x + 1) / 2 - mean) / std is folded
into the first conv (a per-channel shift in the layout conversion, a per-channel factor in its weights).

Weights: the reference loads torchvision's ImageNet checkpoint (vgg19-dcbb9e9d.pth); any state_dict with torchvision's
`features.<idx>.weight|bias` keys (or the reference module's `vgg_net.convX_Y.*`) is accepted.  Offline there is no checkpoint:
tests and bench.py use a seeded random VGG19 — the arithmetic is the same, the weights are data.
"""
import math

import torch

from . import _lib, ops

F16, F32 = torch.float16, torch.float32
VGG19_NAMES = ['conv1_1', 'relu1_1', 'conv1_2', 'relu1_2', 'pool1', 'conv2_1', 'relu2_1', 'conv2_2', 'relu2_2', 'pool2',
               'conv3_1', 'relu3_1', 'conv3_2', 'relu3_2', 'conv3_3', 'relu3_3', 'conv3_4', 'relu3_4', 'pool3', 'conv4_1',
               'relu4_1', 'conv4_2', 'relu4_2', 'conv4_3', 'relu4_3', 'conv4_4', 'relu4_4', 'pool4', 'conv5_1', 'relu5_1',
               'conv5_2', 'relu5_2', 'conv5_3', 'relu5_3', 'conv5_4', 'relu5_4', 'pool5']
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


class VGG19Features:
    """Frozen VGG19 feature extractor packed for the kernels: per conv the forward operand [cout][9*cin] and the adjoint
    operand of its input gradient [cin][9*cout] (b200ir_pack_weights), fp32 bias."""

    def __init__(self, state_dict, layer_names, device, use_input_norm=True, range_norm=False):
        self.layer_names = list(layer_names)
        self.dev = device
        last = max(VGG19_NAMES.index(n) for n in self.layer_names)
        self.program = VGG19_NAMES[:last + 1]
        for i, name in enumerate(self.program):
            if name in self.layer_names and i + 1 < len(self.program):
                nxt = self.program[i + 2] if i + 2 < len(self.program) else None
                if nxt is not None and not nxt.startswith('pool'):
                    raise NotImplementedError(f'tapped layer {name}: only taps at the last conv of a block (followed by ReLU + '
                                              'MaxPool, or the last layer) are implemented — every shipped config uses those')
        # input normalisation (vgg_arch.py:146-149): x_n = (x - sub_c) * m_c
        sub = [0.0, 0.0, 0.0]
        mul = [1.0, 1.0, 1.0]
        if range_norm:
            sub, mul = [-1.0] * 3, [0.5] * 3
        if use_input_norm:
            sub = [s + MEAN[c] / m for c, (s, m) in enumerate(zip(sub, mul))]
            mul = [m / STD[c] for c, m in enumerate(mul)]
        self.sub = torch.tensor(sub, device=device, dtype=F32)
        self.convs = {}
        for idx, name in enumerate(self.program):
            if not name.startswith('conv'):
                continue
            w = self._get(state_dict, idx, name, 'weight').to(device, F32)
            b = self._get(state_dict, idx, name, 'bias').to(device, F32).contiguous()
            if idx == 0:
                w = w * torch.tensor(mul, device=device, dtype=F32).view(1, 3, 1, 1)
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 13))        # 3 -> 16 input channels (zero weights)
            w = w.contiguous()
            self.convs[name] = dict(w=ops.pack_weights(w, 1.0, 0), wt=ops.pack_weights(w, 1.0, 1), b=b, cin=w.shape[1],
                                    cout=w.shape[0])

    @staticmethod
    def _get(sd, idx, name, kind):
        for key in (f'features.{idx}.{kind}', f'vgg_net.{name}.{kind}', f'{idx}.{kind}', f'{name}.{kind}'):
            if key in sd:
                return sd[key].detach()
        raise KeyError(f'VGG19 state_dict has no {kind} for {name} (features.{idx})')

    # ------------------------------------------------------------------ forward
    def forward(self, x, keep):
        """x fp32 NCHW [B,3,H,W] -> ({layer: pre-ReLU feature NHWC fp16}, tape).  keep=False drops everything but the features."""
        B, _, H, W = x.shape
        e16 = lambda *s: torch.empty(*s, device=x.device, dtype=F16)   # noqa: E731
        cur = e16(B, H, W, 16)
        ops.nchw_to_nhwc_pad(x.contiguous(), cur, self.sub, 1.0)
        feats, tape = {}, []
        h, w = H, W
        i = 0
        prog = self.program
        while i < len(prog):
            name = prog[i]
            if name.startswith('conv'):
                c = self.convs[name]
                out = e16(B, h, w, c['cout'])
                tapped = name in self.layer_names
                if tapped:
                    ops.conv_same(cur, c['w'], out, 3, bias=c['b'])()                       # z, read before its ReLU
                    feats[name] = out
                    tape.append(('conv', name, None))
                    i += 1
                    if i < len(prog):                                                         # relu (+ pool, checked in __init__)
                        i += 1
                        if i < len(prog):
                            pooled = e16(B, h // 2, w // 2, c['cout'])
                            ops.maxpool2_relu(out, pooled)
                            tape.append(('pool', name, out))
                            cur, h, w = pooled, h // 2, w // 2
                            i += 1
                    continue
                ops.conv_same(cur, c['w'], out, 3, bias=c['b'], act_slope=0.0)()             # conv + bias + ReLU
                tape.append(('conv', name, out if keep else None))
                cur = out
                i += 2                                                                       # the fused relu
                if i < len(prog) and prog[i].startswith('pool'):
                    pooled = e16(B, h // 2, w // 2, c['cout'])
                    ops.maxpool2_relu(out, pooled)
                    tape.append(('pool', None, out if keep else None))
                    cur, h, w = pooled, h // 2, w // 2
                    i += 1
                continue
            raise AssertionError(name)
        return feats, tape

    # ------------------------------------------------------------------ backward (input gradient only: the VGG is frozen)
    def backward(self, tape, dfeat):
        """dfeat: {layer: d loss / d feature (NHWC fp16)} -> d loss / d x' as NHWC fp16 [B,H,W,16] (channels 0..2)."""
        d = None
        masked = False           # d already carries the ReLU mask of the conv output it belongs to (came out of a pool backward)
        for kind, name, t in reversed(tape):
            if kind == 'pool':
                add = dfeat.get(name) if name is not None else None
                dz = torch.empty_like(t)
                ops.maxpool2_relu_bwd(t, d, add, dz)
                d, masked = dz, True
                continue
            c = self.convs[name]
            if name in dfeat and d is None:                  # the top tap: nothing above it
                d = dfeat[name]
            elif t is not None and not masked:               # conv + ReLU: dz = dy * (y > 0)
                d, _ = ops.lrelu_bias_bwd(d, t, slope=0.0, scale=1.0, want_bias=False)
            b, h, w, _ = d.shape
            dx = torch.empty(b, h, w, c['cin'], device=d.device, dtype=F16)
            ops.conv_same(d, c['wt'], dx, 3)()
            d, masked = dx, False
        return d


def gram_raw(f):
    """sum_p f[b,p,i] f[b,p,j] per image (the un-normalised Gram matrix of losses.py:343-356): NHWC fp16 [B,h,w,C] -> fp32
    [B,C,C], the whole batch in one launch of the weight-gradient GEMM (b200ir_gram_batched)."""
    return ops.gram_batched(f)


class PerceptualLossFunction(torch.autograd.Function):
    """l_g_percep + l_g_style of GFPGANModel.optimize_parameters (gfpgan_model.py:538-545) as ONE differentiable scalar; the
    two values are also returned (detached) for the log.  Forward runs the VGG on x and gt, evaluates both terms and sends
    their gradients back through the VGG right away (input gradient only), so nothing of the VGG outlives the call; backward
    hands out the stored d/dx.  The internal backward runs at grad_scale x 1024: the mean over C x H x W feature elements
    makes these gradients ~1e3 smaller than the pixel loss's."""

    @staticmethod
    def forward(ctx, x, gt, vgg, layer_weights, perceptual_weight, style_weight, grad_scale):
        _lib.require_cuda(x, 'perceptual.PerceptualLossFunction')
        B, _, H, W = x.shape
        dev = x.device
        sv = float(grad_scale) * 1024.0
        fx, tape = vgg.forward(x.detach().float(), keep=True)
        fg, _ = vgg.forward(gt.detach().float(), keep=False)
        percep = torch.zeros(1, device=dev, dtype=F32)
        style = torch.zeros(1, device=dev, dtype=F32)
        want = ctx.needs_input_grad[0]
        dfeat = {}
        for name, lw in layer_weights.items():
            a, b = fx[name], fg[name]
            _, h, w, C = a.shape
            grad = torch.empty_like(a) if want else None
            if perceptual_weight > 0:
                ops.l1_loss_f16(a, b, lw * perceptual_weight, sv, percep, grad)
            elif want:
                grad.zero_()
            if style_weight > 0:
                ga, gb = gram_raw(a), gram_raw(b)
                n = ga.numel()
                wgt = lw * style_weight / (C * h * w)
                sign = torch.empty_like(ga) if want else None
                ops.l1_loss(ga, gb, wgt, n / wgt, style, sign)                        # grad = +-1 exactly
                if want:
                    sm = (sign + sign.transpose(1, 2)).to(F16).contiguous()           # {-2..2}: exact in fp16
                    f = torch.full((B, C), sv * wgt / n, device=dev, dtype=F32)        # per-element factor, applied in the epilogue
                    # dF_b = f * F_b (S_b + S_b^T) + the L1 term: ONE 1x1 conv whose matrix differs per image (w_per_image)
                    ops.conv_same(a, sm, grad, 1, demod=f, res=grad, res_mode=1, res_strides=(C, w * C, h * w * C),
                                  res_wh=(w, h), res_scale=1.0, res_mul=1.0, w_per_image=True)()
            if want:
                dfeat[name] = grad
        dx = None
        if want:
            d16 = vgg.backward(tape, dfeat)
            dx = torch.empty(B, 3, H, W, device=dev, dtype=F32)
            ops.head_to_nchw(d16, dx)
            dx.mul_(1.0 / 1024.0)
        ctx.dx, ctx.grad_scale = dx, float(grad_scale)
        total, p_out, s_out = (percep + style)[0], percep[0], style[0]
        ctx.mark_non_differentiable(p_out, s_out)
        return total, p_out, s_out

    @staticmethod
    def backward(ctx, dl, _dp, _ds):
        dx = ctx.dx
        ctx.dx = None
        if dx is not None:
            dx = dx.mul_(dl / ctx.grad_scale)
        return dx, None, None, None, None, None, None


def perceptual_loss(x, gt, vgg, layer_weights, perceptual_weight=1.0, style_weight=0.0, grad_scale=1.0):
    """Returns (l_g_percep + l_g_style [differentiable w.r.t. x], l_g_percep, l_g_style)."""
    return PerceptualLossFunction.apply(x, gt, vgg, dict(layer_weights), float(perceptual_weight), float(style_weight), grad_scale)
