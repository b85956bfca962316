"""R1 regularisation of the discriminator on the B200 kernels (GFPGANModel.optimize_parameters, gfpgan_model.py:683-689;
r1_penalty, basicsr/losses/losses.py:492-506):

    real_pred = net_d(gt);  grad_real = autograd.grad(real_pred.sum(), gt, create_graph=True)
    l_d_r1 = r1_reg_weight / 2 * mean_b |grad_real_b|^2 * net_d_reg_every;  l_d_r1.backward()

The reference differentiates twice through fused_act / upfirdn2d (ops/fused_act/fused_act.py:30-63, ops/upfirdn2d/upfirdn2d.py:
22-86).  Here the double backward is assembled from first-order passes over the kernels that already exist.  With
f(x; theta) = sum_b D(x)_b, g = grad_x f and P = c / B * |g|^2:

    grad_theta P = 2 c / B * grad_theta [ g(theta) . v ]  at v = g,     g . v = D_v f = directional derivative of f along v,

and D_v f is a forward-mode (tangent) pass: t_0 = v, per layer t_out = phi'(z) * (W t_in) with the primal activation pattern;
its parameter gradient is  dW_l = dz_l (x) t_l  with dz_l = d f / d z_l, the PRIMAL backward signal at the layer's
pre-activation (the adjoint of the tangent network is the primal adjoint).  So:

    1. primal forward of net_d on the real images (every activation kept)
    2. primal backward with d(score) = 1, keeping dz_l of every conv / linear; it ends in g, and P = c / B * sum g^2
    3. tangent forward along g: the same conv / FIR kernels without bias, b200ir_lrelu_bias_bwd as the activation's derivative
    4. dW_l += 2 c / B * wgrad(dz_l, t_l)  -- the weight-gradient kernels on (primal signal, tangent) pairs; biases get nothing
    5. the one layer that is not piecewise linear, the minibatch standard deviation (stylegan2_arch.py:791-801), adds the term
       through its Hessian: q = H[a . s](y) t_y (b200ir_minibatch_stddev_hvp), sent down the layers below it as an ordinary
       backward pass (weights and biases).

All activations NHWC fp16; the backward signals and the tangents carry static scales (s1, s2) that are divided out of the fp32
weight gradients (sq: the Hessian term's own scale).  torch allocates, packs weights and adds the results into .grad.
"""
import math

import torch

from . import _lib, ops
from .backward import pack_equal_conv, pack_equal_conv3x3, resblock_backward, resblock_forward, smoothed_buffer

F16, F32 = torch.float16, torch.float32


def _e16(like, *shape):
    return torch.empty(*shape, device=like.device, dtype=F16)


def r1_penalty_backward(sd, x, weight, grad_out_scale=1.0, stddev_group=4, s1=256.0, s2=256.0, sq=4096.0):
    """l_d_r1 = weight * mean_b |grad_x sum D(x)|^2 for the discriminator `sd` (fp32 CUDA parameters under the reference's
    state_dict names) on the real batch x (fp32 NCHW [B,3,H,W]); weight = r1_reg_weight / 2 * net_d_reg_every.
    Adds grad_out_scale * d(l_d_r1)/d(param) to every parameter's .grad (grad_out_scale = the trainer's loss scale, which the
    fused Adam step divides out) and returns the penalty as a 0-dim fp32 tensor."""
    _lib.require_cuda(x, 'r1.r1_penalty_backward')
    with torch.no_grad():
        return _r1(sd, x.contiguous().float(), float(weight), float(grad_out_scale), stddev_group, float(s1), float(s2), float(sq))


def _add_grad(p, g):
    g = g.reshape(p.shape).to(p.dtype)
    if p.grad is None:
        p.grad = g.clone()
    else:
        p.grad.add_(g)


def _r1(sd, x, weight, gscale, stddev_group, s1, s2, sq):
    dev = x.device
    B, _, H, W = x.shape
    inv = ops.INV_SQRT2
    # ------------------------------------------------------------------ 1. primal forward
    w0 = sd['conv_body.0.0.weight']
    C0 = w0.shape[0]
    w0s = (w0.detach().view(C0, 3) * (1.0 / math.sqrt(3.0))).contiguous()
    y0 = _e16(x, B, H, W, C0)
    ops.first_conv(x, w0s, sd['conv_body.0.1.bias'].detach().float().contiguous(), y0)
    blocks, feat, i = [], y0, 1
    while f'conv_body.{i}.conv1.0.weight' in sd:
        pre = f'conv_body.{i}'
        out, saved, scales = resblock_forward(feat, sd[f'{pre}.conv1.0.weight'], sd[f'{pre}.conv1.1.bias'],
                                              sd[f'{pre}.conv2.1.weight'], sd[f'{pre}.conv2.2.bias'], sd[f'{pre}.skip.1.weight'])
        blocks.append((pre, saved, scales))
        feat = out
        i += 1
    b, h, w, c = feat.shape
    group = min(B, stddev_group)
    M = B // group
    cat = ops.minibatch_stddev(feat, group)
    c_pad = cat.shape[3]
    kf = math.sqrt(c_pad / (c + 1.0))                                   # keeps the 1 / sqrt((C + 1) * 9) of the unpadded conv
    wf = torch.nn.functional.pad(sd['final_conv.0.weight'].detach(), (0, 0, 0, 0, 0, c_pad - (c + 1))) * kf
    wfp, sf = pack_equal_conv3x3(wf)
    gact = _e16(x, b, h, w, wfp.shape[0])
    ops.conv_same(cat, wfp, gact, 3, bias=sd['final_conv.1.bias'].detach().float().contiguous(), act=True)()
    c4 = gact.shape[3]
    wl1 = sd['final_linear.0.weight'].detach()
    n1, k1 = wl1.shape
    wl1n = wl1.view(n1, c4, h * w).permute(0, 2, 1).reshape(n1, k1)     # NCHW flattening -> NHWC order
    sl1 = 1.0 / math.sqrt(k1)
    wp1 = (wl1n * sl1).to(F16).contiguous()
    hid = _e16(x, B, n1)
    ops.conv_same(gact.view(1, 1, B, k1), wp1, hid.view(1, 1, B, n1), 1, bias=sd['final_linear.0.bias'].detach().float().contiguous(),
                  act=True)()
    wl2 = sd['final_linear.1.weight'].detach()                           # [1, n1] -> padded to 16 rows
    sl2 = 1.0 / math.sqrt(n1)
    wp2 = torch.zeros(16, n1, device=dev, dtype=F16)
    wp2[:1] = (wl2 * sl2).to(F16)
    # ------------------------------------------------------------------ 2. primal backward, d(score) = s1
    dscore = torch.zeros(B, 16, device=dev, dtype=F16)
    dscore[:, 0] = s1
    dhid = _e16(x, B, n1)
    ops.conv_same(dscore.view(1, 1, B, 16), wp2.t().contiguous(), dhid.view(1, 1, B, n1), 1)()
    dz_h, _ = ops.lrelu_bias_bwd(dhid, hid, want_bias=False)
    dg = _e16(x, B, k1)
    ops.conv_same(dz_h.view(1, 1, B, n1), wp1.t().contiguous(), dg.view(1, 1, B, k1), 1)()
    dz_f, _ = ops.lrelu_bias_bwd(dg.view(b, h, w, c4), gact, want_bias=False)
    dcat = _e16(x, b, h, w, c_pad)
    ops.conv_dgrad(dz_f, ops.conv_dgrad_weight(wfp, c_pad), dcat)()
    a = dcat[..., c].float().view(group, M, h * w).sum(dim=(0, 2)).contiguous()       # d f / d statistic[m] (B * h * w numbers)
    dfeat = torch.empty_like(feat)
    ops.minibatch_stddev_bwd(feat, dcat, a, dfeat, group)
    dzs = []
    d = dfeat
    need_x = (True, False, False, False, False, False)
    for pre, saved, scales in reversed(blocks):
        res = resblock_backward(saved, scales, d, need_x)
        dzs.append((d, res[6], res[7]))                                  # (dout, dz1, dz2) of this block
        d = res[0]
    dzs.reverse()
    dz0, _ = ops.lrelu_bias_bwd(d, y0, want_bias=False)
    g = torch.empty(B, 3, H, W, device=dev, dtype=F32)                   # = s1 * grad_x f
    ops.first_conv_dgrad(dz0, w0s, g)
    penalty = torch.zeros(1, device=dev, dtype=F32)
    ops.sum_squares(g, weight / (B * s1 * s1), penalty)
    # ------------------------------------------------------------------ 3. tangent forward along v = g
    t0 = g * (s2 / s1)                                                   # fp32 NCHW, carries s2
    t0h = _e16(x, B, H, W, 16)
    ops.nchw_to_nhwc_pad(t0, t0h)
    w0pad = torch.zeros(C0, 16, device=dev, dtype=F16)
    w0pad[:, :3] = w0s.to(F16)
    z = _e16(x, B, H, W, C0)
    ops.conv_same(t0h, w0pad, z, 1)()
    tx, _ = ops.lrelu_bias_bwd(z, y0, dz=z, want_bias=False)             # phi'(z0) * (W0 t0)
    # ------------------------------------------------------------------ 4. weight gradients dz (x) tangent, layer by layer
    k = gscale * 2.0 * weight / (B * s1 * s2)
    _add_grad(w0, ops.first_conv_wgrad(t0, dz0) * (k / math.sqrt(3.0)))
    for (pre, saved, scales), (dout, dz1, dz2) in zip(blocks, dzs):
        xb, t1, p, y2, sk_in, w1p, w2raw, wsp = saved
        sc1, sc2, scs = scales
        bb, hh, ww, cin = xb.shape
        cout = y2.shape[3]
        _add_grad(sd[f'{pre}.conv1.0.weight'], ops.conv_wgrad(tx, dz1).view(cin, 3, 3, cin).permute(0, 3, 1, 2) * (k * sc1))
        tz1 = torch.empty_like(t1)
        ops.conv_same(tx, w1p, tz1, 3)()
        tt1, _ = ops.lrelu_bias_bwd(tz1, t1, dz=tz1, want_bias=False)
        tp = smoothed_buffer(tt1)
        _add_grad(sd[f'{pre}.conv2.1.weight'],
                  ops.conv3x3_s2_wgrad(tp, hh, ww, dz2).view(cout, 3, 3, cin).permute(0, 3, 1, 2) * (k * sc2))
        w2p, _ = pack_equal_conv(w2raw)
        tz2 = torch.empty_like(y2)
        ops.conv3x3_s2(tp, hh, ww, w2p, tz2)()
        ty2, _ = ops.lrelu_bias_bwd(tz2, y2, dz=tz2, want_bias=False)
        tsk = torch.empty_like(sk_in)
        ops.fir_down2(tx, tsk)
        _add_grad(sd[f'{pre}.skip.1.weight'], ops.conv1x1_wgrad(tsk, dout) * (k * scs * inv))
        oh, ow = hh // 2, ww // 2
        tout = _e16(x, bb, oh, ow, cout)
        ops.conv_same(tsk, wsp, tout, 1, res=ty2, res_mode=1, res_strides=(cout, ow * cout, oh * ow * cout), res_wh=(ow, oh),
                      res_scale=inv)()
        tx = tout
    tfeat = tx
    tcat = ops.minibatch_stddev_jvp(feat, tfeat, group)
    dwf = ops.conv_wgrad(tcat, dz_f).view(wfp.shape[0], 3, 3, c_pad).permute(0, 3, 1, 2)[:, :c + 1] * (k * sf * kf)
    _add_grad(sd['final_conv.0.weight'], dwf)
    tzf = torch.empty_like(gact)
    ops.conv_same(tcat, wfp, tzf, 3)()
    tg, _ = ops.lrelu_bias_bwd(tzf, gact, dz=tzf, want_bias=False)
    dwl1 = ops.conv1x1_wgrad(tg.view(1, 1, B, k1), dz_h.view(1, 1, B, n1)) * (k * sl1)            # NHWC column order
    _add_grad(sd['final_linear.0.weight'], dwl1.view(n1, h * w, c4).permute(0, 2, 1))
    tzh = _e16(x, B, n1)
    ops.conv_same(tg.view(1, 1, B, k1), wp1, tzh.view(1, 1, B, n1), 1)()
    thid, _ = ops.lrelu_bias_bwd(tzh, hid, dz=tzh, want_bias=False)
    dwl2 = ops.conv1x1_wgrad(thid.view(1, 1, B, n1), dscore.view(1, 1, B, 16))[:1] * (k * sl2)
    _add_grad(sd['final_linear.1.weight'], dwl2)
    # ------------------------------------------------------------------ 5. Hessian term of the minibatch standard deviation
    # q is small (a / (C P) times a tangent): a third static scale keeps this pass out of the fp16 subnormals on its way down
    q = ops.minibatch_stddev_hvp(feat, tfeat, (a * sq).contiguous(), group)          # carries s1 * s2 * sq
    kq = k / sq
    need_all = (True, True, True, True, True, True)
    d = q
    for pre, saved, scales in reversed(blocks):
        dx, dw1, db1, dw2, db2, dws, _, _ = resblock_backward(saved, scales, d, need_all)
        for name, gr in ((f'{pre}.conv1.0.weight', dw1), (f'{pre}.conv1.1.bias', db1), (f'{pre}.conv2.1.weight', dw2),
                         (f'{pre}.conv2.2.bias', db2), (f'{pre}.skip.1.weight', dws)):
            _add_grad(sd[name], gr * kq)
        d = dx
    dzq, dbq = ops.lrelu_bias_bwd(d, y0)
    _add_grad(sd['conv_body.0.1.bias'], dbq * kq)
    _add_grad(w0, ops.first_conv_wgrad(x, dzq) * (kq / math.sqrt(3.0)))
    return penalty[0]
