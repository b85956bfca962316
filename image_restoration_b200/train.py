"""The training step on the B200 kernels (BASELINE config 5): GFPGANModel.optimize_parameters
(basicsr/models/gfpgan_model.py:494-691) for the plate configs (`fix_decoder` true or false): one net_g update on
l_g_pix + image-pyramid + l_g_gan, the EMA, one net_d update on the logistic loss — data-parallel with one NCCL all-reduce of
the flat gradient buffer per network (base_model.py:62-76).

    train_forward(net, lq)          differentiable GFPGANv1OCR.forward(lq, return_rgb=True): U-Net through backward.unet_forward,
                                    the StyleGAN2 decoder through DecoderFunction
    DecoderFunction                 StyleGAN2OCRGeneratorSFT.forward (gfpganv1_ocr_arch.py:50-136) w.r.t. style code and SFT conditions,
                                    and w.r.t. the decoder's own parameters when they require grad (fix_decoder=False)
    l1_loss / gan_softplus_loss     L1Loss(mean) / GANLoss('wgan_softplus') (losses/losses.py:81-106, 404-470), value + gradient in one pass
    disc_forward_image              network_d on an fp32 NCHW image, with the input gradient the generator needs
    GFPGANTrainer                   feed_data / optimize_parameters / model_ema, mirrors GFPGANModel for these options

Gradients of activations are NHWC fp16: every loss is multiplied by a static loss scale (default 4096 x batch, so that
d(loss)/d(pixel) of a mean-reduced loss lands near 1e-3..1e+2) that FlatAdam's grad_scale removes inside the fused step.
torch runs the autograd graph, allocates buffers and owns the NCCL transport; all arithmetic is in libb200ir.so.
"""
import math
import weakref

import torch

from . import _lib, ops
from .backward import unet_forward
from .engine import pack_decoder

F16, F32 = torch.float16, torch.float32


# ------------------------------------------------------------------------------------------ frozen decoder
class DecoderState:
    """Frozen StyleGAN2 decoder packed for the training path: forward operands + the adjoint weight packs of the input-gradient
    convs.  Built once per (module, decoder parameter versions)."""

    def __init__(self, net):
        # fix_decoder=False (the shipped training YAMLs): the decoder's own parameters get gradients as well
        self.trainable = any(p.requires_grad for p in net.stylegan_decoder.parameters())
        self.param_names = [n for n, _ in net.stylegan_decoder.named_parameters()]
        if not net.input_is_latent:
            raise NotImplementedError('training path: input_is_latent=False (style MLP) is not differentiated')
        dev = next(net.parameters()).device
        self.dev = dev
        sd = {k: v.detach() for k, v in net.state_dict().items() if k.startswith('stylegan_decoder.')}
        pack_decoder(self, net, lambda k: sd[k].to(dev), lambda k: sd[k].to(dev).float().contiguous(), train=True)
        if self.trainable:      # fp32 weights in the reference layout: the demodulation's dependence on W needs them
            D = 'stylegan_decoder'
            self.sc1['w_raw'] = sd[f'{D}.style_conv1.modulated_conv.weight'][0].float()
            for j, c in enumerate(self.sconv):
                c['w_raw'] = sd[f'{D}.style_convs.{j}.modulated_conv.weight'][0].float()
            self.const_raw = sd[f'{D}.constant_input.weight'][0].float()
        self.different_w, self.sft_half, self.nf = net.different_w, net.sft_half, net.num_style_feat
        self.H, self.W = net.input_height, net.input_width
        self.ratio = int(self.W / self.H)
        self.sig = _decoder_signature(net)


_DECODERS = weakref.WeakKeyDictionary()


def _decoder_signature(net):
    ps = list(net.stylegan_decoder.parameters()) + list(net.stylegan_decoder.buffers())
    return (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))


def decoder_state(net):
    st = _DECODERS.get(net)
    if st is None or st.sig != _decoder_signature(net):
        with torch.no_grad():
            st = DecoderState(net)
        _DECODERS[net] = st
    return st


def _convt_merged(cout, B, h, w):
    """Same choice as the inference plan (engine._Plan): merged form for the 64/128-channel levels and the small ones."""
    return cout <= 128 or B * h * w < 40000


class DecoderFunction(torch.autograd.Function):
    """image = stylegan_decoder([style_code], conditions) (gfpganv1_ocr_arch.py:387-391, 50-136) with gradients for the style
    code and the SFT conditions, and — when the decoder is trainable (fix_decoder=False, the shipped training YAMLs) — for
    every decoder parameter on the path (the style MLP is unused with input_is_latent=True: no gradient, as in the reference).

    forward : existing forward kernels, but every StyleConv keeps its un-modulated output `a` (post-activation, pre-SFT) —
              the inference plan fuses SFT and the next modulation into the producer and keeps neither;
    backward: per StyleConv  to_rgb_bwd -> sft_mod_bwd (direct style gradient + SFT adjoint) -> style_act_bwd (lrelu, demod
              reduction) -> dgrad on b200ir_conv_igemm with adjoint weights (stride 1) / fir_pad22 + the stride-2 conv over phase
              views (adjoint of conv_transpose2d + FIR), then demod_bwd + mod_linear_bwd into d(latent)."""

    @staticmethod
    def forward(ctx, st, noises, n_conds, style_code, *rest):
        """rest = the n_conds SFT condition tensors, then (fix_decoder=False) the decoder's parameters in named_parameters()
        order — passed only so that autograd routes their gradients; the arithmetic uses the packed copies in `st`."""
        _lib.require_cuda(style_code, 'train.FrozenDecoderFunction')
        conds = rest[:n_conds]
        ctx.n_conds, ctx.n_params = n_conds, len(rest) - n_conds
        keep = st.trainable          # the weight gradients need every conv's modulated input
        dev = style_code.device
        B = style_code.shape[0]
        L = st.L
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=F16)   # noqa: E731
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=F32)   # noqa: E731
        latent = style_code.detach().float().reshape(B, -1, st.nf).contiguous()       # [B, num_latent | 1, F]
        lat = (lambda i: i) if st.different_w else (lambda i: 0)
        mod_layers, demod_layers = [], []

        def mod(layer, li):
            s = e32(B, layer['mod_w'].shape[0])
            mod_layers.append((layer['mod_w'], layer['mod_b'], lat(li), s))
            return s

        def dem(layer, s):
            d = e32(B, layer['cout'])
            demod_layers.append((s, layer['wsq'], layer['scale2'], d))
            return d

        s_sc1 = mod(st.sc1, 0)
        d_sc1 = dem(st.sc1, s_sc1)
        s_rgb1 = mod(st.rgb1, 1)
        s_conv, d_conv, s_rgb = [], [], []
        for lvl in range(L):
            i = 1 + 2 * lvl
            s1, s2 = mod(st.sconv[2 * lvl], i), mod(st.sconv[2 * lvl + 1], i + 1)
            s_conv += [s1, s2]
            d_conv += [dem(st.sconv[2 * lvl], s1), dem(st.sconv[2 * lvl + 1], s2)]
            s_rgb.append(mod(st.rgbs[lvl], i + 2))
        ops.ModLinearMulti(latent, mod_layers, st.mod_wscale)()
        ops.DemodMulti(demod_layers)()

        nz = [n.expand(B, 1, n.shape[2], n.shape[3]).contiguous() if n.shape[0] != B else n.contiguous() for n in noises]
        h, w = 4, 4 * st.ratio
        xs = e16(B, h, w, st.sc1['cin'])
        ops.modulate_const(st.const, s_sc1, xs)
        xs_in0 = xs if keep else None
        a0 = e16(B, h, w, st.sc1['cout'])
        ops.conv_same(xs, st.sc1['w'], a0, 3, bias=st.sc1['bias'], demod=d_sc1, noise=nz[0], noise_gain=st.sc1['gain'],
                      noise_strides=(h * w, w), act=True)()
        skip = e32(B, 3, h, w)
        xs = e16(B, h, w, st.sc1['cout']) if L else None
        ops.to_rgb(a0, st.rgb1['w'], s_rgb1, st.rgb1['bias'], None, skip, s_next=s_conv[0] if L else None, xs_out=xs)
        acts = []
        for lvl in range(L):
            c1, c2 = st.sconv[2 * lvl], st.sconv[2 * lvl + 1]
            cout = c1['cout']
            h2, w2 = 2 * h, 2 * w
            raw = e16(B, h2 + 2, w2 + 2, cout)       # valid (2h+1) x (2w+1); the spare row / column is never read
            if _convt_merged(cout, B, h, w):
                ops.convt_s2_merged(xs, c1['w_merged'], raw, d_conv[2 * lvl])()
            else:
                for pi, (py, px) in enumerate(ops.CONVT_PHASES):
                    ops.convt_s2_phase(xs, c1['w_phase'][pi], py, px, raw, demod=d_conv[2 * lvl])()
            xs_up = xs if keep else None
            a1 = e16(B, h2, w2, cout)
            ops.upfir_act(raw, a1, nz[2 * lvl + 1], h2 * w2, c1['gain'], c1['bias'], None, None, 0, None)
            del raw
            sc, sh = conds[2 * lvl].detach(), conds[2 * lvl + 1].detach()
            xs2 = e16(B, h2, w2, cout)
            ops.sft_mod(a1, sc, sh, s_conv[2 * lvl + 1], xs2)
            a2 = e16(B, h2, w2, c2['cout'])
            ops.conv_same(xs2, c2['w'], a2, 3, bias=c2['bias'], demod=d_conv[2 * lvl + 1], noise=nz[2 * lvl + 2],
                          noise_gain=c2['gain'], noise_strides=(h2 * w2, w2), act=True)()
            xs_keep = xs2 if keep else None
            del xs2
            last = lvl == L - 1
            nskip = e32(B, 3, h2, w2)
            xs = None if last else e16(B, h2, w2, c2['cout'])
            ops.to_rgb(a2, st.rgbs[lvl]['w'], s_rgb[lvl], st.rgbs[lvl]['bias'], skip, nskip,
                       s_next=None if last else s_conv[2 * lvl + 2], xs_out=xs)
            skip = nskip
            acts.append((a1, a2, sc, sh, xs_up, xs_keep))
            h, w = h2, w2
        ctx.st, ctx.nz, ctx.acts, ctx.a0, ctx.xs_in0 = st, nz, acts, a0, xs_in0
        ctx.latent = latent if keep else None
        ctx.tables = (s_sc1, d_sc1, s_rgb1, s_conv, d_conv, s_rgb)
        ctx.lat_shape = latent.shape
        ctx.code_shape, ctx.code_dtype = style_code.shape, style_code.dtype
        return skip

    @staticmethod
    def backward(ctx, d_image):
        st, nz, acts, a0 = ctx.st, ctx.nz, ctx.acts, ctx.a0
        s_sc1, d_sc1, s_rgb1, s_conv, d_conv, s_rgb = ctx.tables
        dev = d_image.device
        B = d_image.shape[0]
        L = st.L
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=F16)   # noqa: E731
        z32 = lambda *s: torch.zeros(*s, device=dev, dtype=F32)   # noqa: E731
        lat = (lambda i: i) if st.different_w else (lambda i: 0)
        dlat = z32(*ctx.lat_shape)
        wscale = st.mod_wscale
        dskip = d_image.contiguous().float()
        d_conds = [None] * (2 * L)
        g_next = None           # gradient w.r.t. the modulated input of the conv that consumed this level's a2 (next level's conv1)
        nxt = None              # (layer, s, d, dd, lat index, name) of that conv: its style gradient needs this level's a2
        train_p = st.trainable  # fix_decoder=False: parameter gradients as well, keyed by the decoder's parameter names
        pg = {}
        latent = ctx.latent

        def mod_param_grads(name, ds, li):
            """modulation EqualLinear of a ModulatedConv2d (stylegan2_ocr_arch.py:233-234, lr_mul 1): weight and bias."""
            pg[f'{name}.modulated_conv.modulation.weight'] = ops.mod_linear_wgrad(ds, latent, wscale, lat(li))
            pg[f'{name}.modulated_conv.modulation.bias'] = ops.table_colsum(ds)

        def style_grad(layer, s, d, dd, ds, li, name):
            ops.demod_bwd(ds, s, dd, d, layer['wsq'], layer['scale2'])
            ops.mod_linear_bwd(ds, layer['mod_w'], wscale, dlat, lat(li))
            if train_p:
                mod_param_grads(name, ds, li)

        def act_bwd(name, layer, da, a, noise, d, mul, dd):
            """noise + FusedLeakyReLU backward in place (da becomes dy * d * mul); with parameters also activate.bias and the
            noise gain StyleConv.weight (stylegan2_ocr_arch.py:316-333)."""
            if not train_p:
                ops.style_act_bwd(da, a, noise, layer['gain'], layer['bias'], d, mul, da, dd)
                return
            C = a.shape[3]
            db, dn = z32(B, C), z32(B, C)
            ops.style_act_bwd_params(da, a, noise, layer['gain'], layer['bias'], d, mul, da, dd, db, dn)
            pg[f'{name}.activate.bias'] = ops.table_colsum(db)
            pg[f'{name}.weight'] = ops.table_colsum(ops.table_colsum(dn).view(C, 1))

        def conv_weight_grad(name, layer, G, transposed, s, dd, d):
            pg[f'{name}.modulated_conv.weight'] = ops.modconv_wgrad(G, transposed, layer['w_raw'], s, dd, d,
                                                                    math.sqrt(layer['scale2'])).unsqueeze(0)

        def rgb_bwd(name, layer, a, s, da, li):
            """ToRGB (modulated 1x1 conv without demodulation + bias, stylegan2_ocr_arch.py:357-374) at the current dskip."""
            C = a.shape[3]
            ds = z32(B, C)
            if train_p:
                R = z32(B, 3, C)
                ops.to_rgb_bwd_params(dskip, a, layer['w'], s, da, False, ds, R)
                pg[f'{name}.modulated_conv.weight'] = ops.table_colsum(R, mul=s, scale=1.0 / math.sqrt(C)).view(1, 3, C, 1, 1)
                bias = z32(3)
                ops.plane_sums(dskip, bias)
                pg[f'{name}.bias'] = bias.view(1, 3, 1, 1)
                mod_param_grads(name, ds, li)
            else:
                ops.to_rgb_bwd(dskip, a, layer['w'], s, da, False, ds)
            ops.mod_linear_bwd(ds, layer['mod_w'], wscale, dlat, lat(li))

        for lvl in range(L - 1, -1, -1):
            a1, a2, sc, sh, xs_up, xs2 = acts[lvl]
            c1, c2 = st.sconv[2 * lvl], st.sconv[2 * lvl + 1]
            n1, n2 = f'style_convs.{2 * lvl}', f'style_convs.{2 * lvl + 1}'
            i = 1 + 2 * lvl
            _, h2, w2, C = a2.shape
            # ---- ToRGB (modulated 1x1, no demod) + skip up-sampling
            da2 = e16(B, h2, w2, C)
            rgb_bwd(f'to_rgbs.{lvl}', st.rgbs[lvl], a2, s_rgb[lvl], da2, i + 2)
            dprev = torch.empty(B, 3, h2 // 2, w2 // 2, device=dev, dtype=F32)
            ops.rgb_up_adjoint(dskip, dprev)
            dskip = dprev
            # ---- the next level's conv1 read a2 * s
            if g_next is not None:
                layer, s, d, dd, li, name = nxt
                ds = z32(B, C)
                ops.sft_mod_bwd(g_next, a2, None, None, s, da2, True, None, None, ds)
                style_grad(layer, s, d, dd, ds, li, name)
            # ---- conv2: a2 = act(conv(xs2) * d + noise + bias)
            dd2 = z32(B, C)
            act_bwd(n2, c2, da2, a2, nz[2 * lvl + 2], d_conv[2 * lvl + 1], 1.0, dd2)
            if train_p:
                conv_weight_grad(n2, c2, ops.conv_wgrad(xs2, da2), False, s_conv[2 * lvl + 1], dd2, d_conv[2 * lvl + 1])
            g2 = e16(B, h2, w2, c2['cin'])
            ops.conv_same(da2, c2['w_dgrad'], g2, 3)()
            del da2, xs2
            # ---- SFT + modulation between conv1 and conv2
            C1 = a1.shape[3]
            da1 = e16(B, h2, w2, C1)
            dsc, dsh = torch.empty_like(sc), torch.empty_like(sh)
            ds = z32(B, C1)
            ops.sft_mod_bwd(g2, a1, sc, sh, s_conv[2 * lvl + 1], da1, False, dsc, dsh, ds)
            del g2
            style_grad(c2, s_conv[2 * lvl + 1], d_conv[2 * lvl + 1], dd2, ds, i + 1, n2)
            d_conds[2 * lvl], d_conds[2 * lvl + 1] = dsc, dsh
            # ---- conv1 (up-sampling): a1 = act(FIR4(convT(xs) * d) + noise + bias)
            dd1 = z32(B, C1)
            act_bwd(n1, c1, da1, a1, nz[2 * lvl + 1], d_conv[2 * lvl], 4.0, dd1)
            draw = e16(B, h2 + 2, w2 + 2, C1)        # fir_pad22 fills the (h2+1) x (w2+1) region the stride-2 conv reads
            if train_p:                              # the weight gradient contracts over whole pixel tiles: spare row / column finite
                draw[:, h2 + 1].zero_()
                draw[:, :, w2 + 1].zero_()
            ops.fir_pad22(da1, draw)
            del da1
            if train_p:   # conv_transpose2d's weight gradient = the stride-2 conv's with input and output exchanged: [cin][9][cout]
                conv_weight_grad(n1, c1, ops.conv3x3_s2_wgrad(draw, h2, w2, xs_up), True, s_conv[2 * lvl], dd1, d_conv[2 * lvl])
            g_next = e16(B, h2 // 2, w2 // 2, c1['cin'])
            ops.conv3x3_s2(draw, h2, w2, c1['w_dgrad'], g_next)()
            del draw
            nxt = (c1, s_conv[2 * lvl], d_conv[2 * lvl], dd1, i, n1)
        # ---- style_conv1 + to_rgb1 on the constant input
        _, h, w, C = a0.shape
        da0 = e16(B, h, w, C)
        rgb_bwd('to_rgb1', st.rgb1, a0, s_rgb1, da0, 1)
        if g_next is not None:
            layer, s, d, dd, li, name = nxt
            ds = z32(B, C)
            ops.sft_mod_bwd(g_next, a0, None, None, s, da0, True, None, None, ds)
            style_grad(layer, s, d, dd, ds, li, name)
        dd0 = z32(B, C)
        act_bwd('style_conv1', st.sc1, da0, a0, nz[0], d_sc1, 1.0, dd0)
        if train_p:
            conv_weight_grad('style_conv1', st.sc1, ops.conv_wgrad(ctx.xs_in0, da0), False, s_sc1, dd0, d_sc1)
        g0 = e16(B, h, w, st.sc1['cin'])
        ops.conv_same(da0, st.sc1['w_dgrad'], g0, 3)()
        ds = z32(B, st.sc1['cin'])
        ops.sft_mod_bwd(g0, st.const, None, None, None, None, False, None, None, ds, a_broadcast=True)
        style_grad(st.sc1, s_sc1, d_sc1, dd0, ds, 0, 'style_conv1')
        if train_p:       # ConstantInput (stylegan2_ocr_arch.py:287-301): x' = const * s, so d(const)[c][p] = sum_b s[b][c] g0[b][p][c]
            cin = st.sc1['cin']
            pg['constant_input.weight'] = ops.table_colsum(g0, mul=s_sc1).view(h, w, cin).permute(2, 0, 1).unsqueeze(0).contiguous()
        d_code = dlat.reshape(ctx.code_shape).to(ctx.code_dtype)
        d_params = tuple(pg.get(n) for n in st.param_names[:ctx.n_params]) if ctx.n_params else ()
        return (None, None, None, d_code) + tuple(d_conds) + d_params


FrozenDecoderFunction = DecoderFunction      # the name of the fix_decoder=True-only version


def decoder_apply(net, st, noises, style_code, conds):
    """stylegan_decoder([style_code], conds) through DecoderFunction; with fix_decoder=False the decoder's parameters are
    handed to autograd as inputs (in named_parameters() order) so that their gradients land in .grad."""
    params = [p for _, p in net.stylegan_decoder.named_parameters()] if st.trainable else []
    return DecoderFunction.apply(st, noises, len(conds), style_code, *conds, *params)


# ------------------------------------------------------------------------------------------ heads / losses
class HeadToNchwFunction(torch.autograd.Function):
    """toRGB head [B,h,w,16] fp16 (backward.ToRGBHeadFunction) -> fp32 NCHW [B,3,h,w], what the reference returns in out_rgbs."""

    @staticmethod
    def forward(ctx, head):
        _lib.require_cuda(head, 'train.HeadToNchwFunction')
        b, h, w, cpad = head.shape
        rgb = torch.empty(b, 3, h, w, device=head.device, dtype=F32)
        ops.head_to_nchw(head.contiguous(), rgb)
        ctx.shape = head.shape
        return rgb

    @staticmethod
    def backward(ctx, drgb):
        dhead = torch.empty(ctx.shape, device=drgb.device, dtype=F16)
        ops.nchw_to_head(drgb.contiguous().float(), dhead)
        return dhead


class L1LossFunction(torch.autograd.Function):
    """weight * mean|x - t| (L1Loss, losses.py:81-106); value and gradient come out of one pass over x and t.  The
    gradient is stored scaled by `grad_scale` (the loss scale); backward multiplies by the incoming scalar / grad_scale."""

    @staticmethod
    def forward(ctx, x, t, weight, grad_scale):
        _lib.require_cuda(x, 'train.L1LossFunction')
        x, t = x.contiguous(), t.contiguous()
        loss = torch.zeros(1, device=x.device, dtype=F32)
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ops.l1_loss(x, t, weight, grad_scale, loss, grad)
        ctx.grad, ctx.grad_scale = grad, grad_scale
        return loss[0]

    @staticmethod
    def backward(ctx, dl):
        g = ctx.grad
        ctx.grad = None
        return _scaled(g, dl, ctx.grad_scale), None, None, None


class SoftplusLossFunction(torch.autograd.Function):
    """weight * mean softplus(sign * pred) (GANLoss 'wgan_softplus', losses.py:404-419): sign = -1 for target_is_real."""

    @staticmethod
    def forward(ctx, pred, sign, weight, grad_scale):
        _lib.require_cuda(pred, 'train.SoftplusLossFunction')
        assert pred.dtype == F16 and pred.dim() == 2 and pred.shape[1] == 1
        pred = pred.contiguous()        # the scores are column 0 of a 16-wide GEMM output: pred and dpred share one stride
        loss = torch.zeros(1, device=pred.device, dtype=F32)
        dpred = torch.empty(pred.shape, device=pred.device, dtype=F16) if ctx.needs_input_grad[0] else None
        ops.softplus_loss(pred, sign, weight, grad_scale, loss, dpred)
        ctx.grad, ctx.grad_scale = dpred, grad_scale
        return loss[0]

    @staticmethod
    def backward(ctx, dl):
        g = ctx.grad
        ctx.grad = None
        return _scaled(g, dl, ctx.grad_scale), None, None, None


def _scaled(g, dl, grad_scale):
    """The stored gradient already carries grad_scale (the loss scale); the incoming scalar dl of
    total.backward(gradient=loss_scale) equals it, so the factor is 1 — applied on the device, no host sync."""
    if g is None:
        return None
    return g.mul_((dl / grad_scale).to(g.dtype))


def l1_loss(x, t, weight=1.0, grad_scale=1.0):
    return L1LossFunction.apply(x, t, weight, grad_scale)


def gan_softplus_loss(pred, target_is_real, weight=1.0, grad_scale=1.0):
    return SoftplusLossFunction.apply(pred, -1.0 if target_is_real else 1.0, weight, grad_scale)


# ------------------------------------------------------------------------------------------ networks
def train_forward(net, x, return_rgb=True, randomize_noise=True, noise=None):
    """Differentiable GFPGANv1OCR.forward (gfpganv1_ocr_arch.py:341-393) for training: returns
    (image fp32 NCHW [B,3,H,W], out_rgbs: list of fp32 NCHW toRGB heads) with autograd history on the trainable parameters.
    noise: optional list of 2L+1 tensors [B|1,1,h,w]; else fresh N(0,1) planes when randomize_noise (the reference's training
    default, gfpgan_model.py:508) or the registered noise buffers."""
    _lib.require_cuda(x, 'train.train_forward')
    st = decoder_state(net)
    sd = dict(net.named_parameters())
    res = unet_forward(sd, x, different_w=net.different_w, num_style_feat=net.num_style_feat, return_rgb=return_rgb)
    style_code, conds = res[0], res[1]
    if noise is None:
        if randomize_noise:
            B = x.shape[0]
            noise = [torch.randn(B, 1, n.shape[2], n.shape[3], device=x.device, dtype=F32) for n in st.stored_noise]
        else:
            noise = st.stored_noise
    image = decoder_apply(net, st, noise, style_code, conds)
    out_rgbs = [HeadToNchwFunction.apply(r) for r in res[2]] if return_rgb else []
    return image, out_rgbs


def disc_forward_image(sd, x, stddev_group=4):
    """network_d on an fp32 NCHW image (backward.disc_forward).  When x requires grad the first conv also returns
    d(score)/d(image): that is what l_g_gan sends back into net_g (gfpgan_model.py:549-552)."""
    from .backward import disc_forward
    return disc_forward(sd, x, stddev_group)


def construct_img_pyramid(gt, levels):
    """GFPGANModel.construct_img_pyramid (gfpgan_model.py:326-332): bilinear x0.5 chain of the ground truth, small to large.
    Data preparation of the targets (no gradient flows into it)."""
    import torch.nn.functional as F
    out = [gt]
    for _ in range(levels - 1):
        out.insert(0, F.interpolate(out[0], scale_factor=0.5, mode='bilinear', align_corners=False))
    return out


class GFPGANTrainer:
    """optimize_parameters of GFPGANModel (gfpgan_model.py:494-691) for the plate options: pixel L1 (weight 0.1), image
    pyramid L1 (weight 1), GAN 'wgan_softplus' (weight 0.1), net_d logistic loss; Adam lr 2e-3 betas (0, 0.99) for both
    networks; EMA decay 0.5 ** (32 / 10000); R1 penalty on the real batch every net_d_reg_every iterations (r1.py); the
    perceptual + style terms when a VGG19 is supplied (perceptual.py).  The identity / component terms are not part of this
    step: the reference never executes them (gfpgan_model.py:70-74 hard-codes use_facial_disc = False and get_roi_regions is
    `pass`; network_identity is commented out in every training YAML).

    net_g: image_restoration_b200.GFPGANv1OCR (fix_decoder True or False) on the device; net_d: image_restoration_b200.disc.StyleGAN2Discriminator;
    net_g_ema: optional second GFPGANv1OCR that receives the EMA of the trainable parameters."""

    def __init__(self, net_g, net_d, net_g_ema=None, lr_g=2e-3, lr_d=2e-3, betas=(0.0, 0.99), pix_weight=0.1, pyramid_weight=1.0,
                 gan_weight=0.1, ema_decay=0.5 ** (32 / (10 * 1000)), loss_scale=None, group=None, net_d_iters=1,
                 net_d_init_iters=0, r1_reg_weight=10.0, net_d_reg_every=16, perceptual=None):
        """perceptual: None, or dict(vgg=perceptual.VGG19Features, layer_weights={...}, perceptual_weight=1.0, style_weight=50.0)
        — the `perceptual_opt` of the training YAMLs (gfpgan_model.py:538-545)."""
        from .grad_sync import GradAllReducer
        from .optim import FlatAdam
        import torch.distributed as dist
        self.net_g, self.net_d, self.net_g_ema = net_g, net_d, net_g_ema
        self.pix_weight, self.pyramid_weight, self.gan_weight = pix_weight, pyramid_weight, gan_weight
        self.ema_decay, self.loss_scale = ema_decay, loss_scale
        self.net_d_iters, self.net_d_init_iters = net_d_iters, net_d_init_iters
        self.r1_reg_weight, self.net_d_reg_every = r1_reg_weight, net_d_reg_every
        self.perceptual = perceptual
        # fp16 activation gradients under a static loss scale can overflow when the losses grow; b200ir_adam_step drops
        # non-finite gradient elements, and every `check_finite_every` iterations the flat gradients are inspected (one host
        # sync): if any element was non-finite the loss scale is halved (never raised again: the default leaves ~2^5 headroom)
        self.check_finite_every = 50
        self.overflow_events = 0
        self.g_params = [p for p in net_g.parameters() if p.requires_grad]
        self.d_params = list(net_d.parameters())
        ema_params = None
        if net_g_ema is not None:
            by_name = dict(net_g_ema.named_parameters())
            ema_params = [by_name[n] for n, p in net_g.named_parameters() if p.requires_grad]
        self.opt_g = FlatAdam(self.g_params, lr=lr_g, betas=betas, ema_params=ema_params)
        self.opt_d = FlatAdam(self.d_params, lr=lr_d, betas=betas)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.sync_g = GradAllReducer(self.g_params, group=group) if self.world > 1 else None
        self.sync_d = GradAllReducer(self.d_params, group=group) if self.world > 1 else None
        self.d_sd = dict(net_d.named_parameters())
        self.log = {}
        self.profile = False        # True: CUDA events at the phase boundaries of optimize_parameters (self.phase_ms())
        self._marks = []

    def _mark(self, name):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def phase_ms(self):
        """Milliseconds between the phase marks of the last optimize_parameters call (profile=True; synchronises)."""
        torch.cuda.synchronize()
        m = self._marks
        return {m[i + 1][0]: m[i][1].elapsed_time(m[i + 1][1]) for i in range(len(m) - 1)}

    def feed_data(self, lq, gt):
        """lq, gt: fp32 NCHW [B,3,H,W] in [-1, 1] on the device (FFHQDegradationDataset pairs; degradation.synthesize_pairs)."""
        self.lq, self.gt = lq.contiguous(), gt.contiguous()

    def _scale(self, B):
        return float(self.loss_scale) if self.loss_scale else 4096.0 * B

    def _backward(self, total, S):
        total.backward(gradient=torch.full_like(total, S))

    def _step(self, opt, sync, S, ema_decay=None):
        """All-reduce (sum) of the flat gradient buffer, then the fused Adam (+ EMA) step; the loss scale and the 1 / world
        average are folded into the step's grad_scale."""
        if sync is not None:
            opt.step(flat_grad=sync.reduce_flat(), grad_scale=1.0 / (S * self.world), ema_decay=ema_decay)
        else:
            opt.step(grad_scale=1.0 / S, ema_decay=ema_decay)

    def optimize_parameters(self, current_iter=1):
        """One iteration of gfpgan_model.py:494-691.  net_d is evaluated on the generator's output ONCE: the reference calls
        net_d(self.output) for l_g_gan and net_d(self.output.detach()) for l_d with identical weights (optimizer_d steps after
        both), so the second forward recomputes the first.  Here the graph of that forward is walked twice — input gradient
        only for the generator (backward.skip_param_grads), parameter gradients only for the discriminator — and the
        generator's own graph is cut at `output` so that it can be freed after its backward."""
        from .backward import graph_tag, skip_param_grads
        lq, gt = self.lq, self.gt
        B = lq.shape[0]
        S = self._scale(B)
        log = {}
        self._marks = []
        self._mark('start')
        # ---------------- optimize net_g (gfpgan_model.py:497-667)
        self.opt_g.zero_grad()
        self.opt_d.zero_grad()
        output, out_rgbs = train_forward(self.net_g, lq, return_rgb=self.pyramid_weight > 0)
        self.output = output.detach()
        self._mark('g_forward')
        fake_pred = None
        g_update = current_iter % self.net_d_iters == 0 and current_iter > self.net_d_init_iters
        if g_update:
            total = l1_loss(output, gt, self.pix_weight, S)                                        # :519-523
            log['l_g_pix'] = total.detach()
            if self.pyramid_weight > 0:                                                             # :531-536
                pyramid_gt = construct_img_pyramid(gt, len(out_rgbs))
                for i, (rgb, tgt) in enumerate(zip(out_rgbs, pyramid_gt)):
                    l_p = l1_loss(rgb, tgt, self.pyramid_weight, S)
                    log[f'l_p_{2 ** (i + 3)}'] = l_p.detach()
                    total = total + l_p
            if self.perceptual is not None:                                                         # :538-545
                from .perceptual import perceptual_loss
                pc = self.perceptual
                l_per, lp, ls = perceptual_loss(output, gt, pc['vgg'], pc['layer_weights'], pc.get('perceptual_weight', 1.0),
                                                pc.get('style_weight', 0.0), S)
                log['l_g_percep'], log['l_g_style'] = lp.detach(), ls.detach()
                total = total + l_per
            out_leaf = self.output.requires_grad_()                                                 # cut: net_d's graph starts here
            with graph_tag('net_d'):
                fake_pred = disc_forward_image(self.d_sd, out_leaf)                                # :549-552
            l_g_gan = gan_softplus_loss(fake_pred, True, self.gan_weight, S)
            log['l_g_gan'] = l_g_gan.detach()
            self._mark('g_losses_and_d_forward')
            with skip_param_grads('net_d'):                                                         # d l_g_gan / d output only
                l_g_gan.backward(gradient=torch.full_like(l_g_gan, S), retain_graph=True)
            d_out, out_leaf.grad = out_leaf.grad, None
            torch.autograd.backward([total, output], [torch.full_like(total, S), d_out])
            self.output = self.output.detach()
            del total, output, out_rgbs, d_out
            self._mark('g_backward')
            self._step(self.opt_g, self.sync_g, S, ema_decay=self.ema_decay if self.net_g_ema is not None else None)
        else:
            del output, out_rgbs
        self._mark('g_allreduce_adam_ema')
        # ---------------- optimize net_d (:672-691)
        if fake_pred is None:
            fake_pred = disc_forward_image(self.d_sd, self.output)
        real_d_pred = disc_forward_image(self.d_sd, gt)
        l_d_real = gan_softplus_loss(real_d_pred, True, 1.0, S)
        l_d_fake = gan_softplus_loss(fake_pred, False, 1.0, S)
        l_d = l_d_real + l_d_fake
        log['l_d'] = l_d.detach()
        log['real_score'] = real_d_pred.detach().float().mean()
        log['fake_score'] = fake_pred.detach().float().mean()
        self._mark('d_forward')
        torch.autograd.backward([l_d], [torch.full_like(l_d, S)], inputs=self.d_params)
        del fake_pred, real_d_pred, l_d, l_d_real, l_d_fake
        self._mark('d_backward')
        if self.r1_reg_weight > 0 and self.net_d_reg_every > 0 and current_iter % self.net_d_reg_every == 0:   # :683-689
            from .r1 import r1_penalty_backward
            log['l_d_r1'] = r1_penalty_backward(self.d_sd, gt, self.r1_reg_weight / 2 * self.net_d_reg_every, grad_out_scale=S,
                                                stddev_group=getattr(self.net_d, 'stddev_group', 4))
            self._mark('d_r1')
        self._step(self.opt_d, self.sync_d, S)
        self._mark('d_allreduce_adam')
        if self.check_finite_every and current_iter % self.check_finite_every == 0:
            flats = [self.sync_g.flat if self.sync_g is not None else self.opt_g.grad,
                     self.sync_d.flat if self.sync_d is not None else self.opt_d.grad]
            if not all(bool(torch.isfinite(f).all()) for f in flats):
                self.overflow_events += 1
                self.loss_scale = self._scale(B) / 2.0
                log['loss_scale_halved_to'] = torch.tensor(self.loss_scale)
        self.log = log
        return log
