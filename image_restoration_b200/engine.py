"""Execution engine of the B200 GFPGANv1OCR: weight pre-packing + per-batch-size launch plans over libb200ir.so.

Data layout in HBM: every activation is NHWC fp16 ([B][H][W][C], C contiguous) so that a 3x3 tap of a 128-position
tile is one TMA box; weights are packed once as fp16 [Cout][taps*Cin] (K contiguous) with the equalised-lr scale
folded in; style vectors, demodulation tables, biases, noise and the RGB skip chain stay fp32.

Follows the data flow of GFPGANv1OCR.forward (gfpganv1_ocr_arch.py:341-393) and
StyleGAN2OCRGeneratorSFT.forward (:50-136); identities used (verified in tests):
  * modulated conv == shared-weight conv of (x * s[b,:]) scaled by demod[b,o] in the epilogue;
  * conv_transpose2d(stride 2) == 4 output-phase GEMMs with 4/2/2/1 taps;
  * bilinear x2 commutes with the 1x1 skip conv of ResUpBlock; FIR-then-sample commutes with the 1x1 skip of ResBlock.
"""
import math
import os

import torch

from . import ops

# SFT heads with at most this many channels run as two chains of two convs (0: always one merged hidden tensor)
_SFT_SPLIT = int(os.environ.get('B200IR_SFT_SPLIT', '64'))

F16 = torch.float16
F32 = torch.float32


def _pack_conv(w, scale):
    """(cout, cin, kh, kw) fp32 -> fp16 [cout][(kh*kw)*cin], scale folded."""
    co, ci, kh, kw = w.shape
    return (w.float() * scale).permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(F16)


class _Packed:
    """Weights of one module instance, packed for the kernels (built once per parameter version)."""

    def __init__(self, net):
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        dev = next(net.parameters()).device
        self.dev = dev
        L = net.log_size - 2
        self.L = L
        g = lambda k: sd[k].to(dev)  # noqa: E731

        def eq(wkey):
            w = g(wkey)
            return _pack_conv(w, 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3]))

        def f32(key):
            return g(key).float().contiguous()

        # encoder
        w0 = g('conv_body_first.0.weight').float()
        self.first_w = (w0[:, :, 0, 0] / math.sqrt(3.0)).contiguous()
        self.first_b = f32('conv_body_first.1.bias')
        self.down = []
        for i in range(L):
            p = f'conv_body_down.{i}'
            self.down.append(dict(w1=eq(f'{p}.conv1.0.weight'), b1=f32(f'{p}.conv1.1.bias'),
                                  w2=eq(f'{p}.conv2.1.weight'), b2=f32(f'{p}.conv2.2.bias'),
                                  ws=eq(f'{p}.skip.1.weight')))
        self.final_w = eq('final_conv.0.weight')
        self.final_b = f32('final_conv.1.bias')
        # final_linear: reference flattens NCHW (c*P + p); our activations are NHWC (p*C + c)
        wl = g('final_linear.weight').float()
        n_out, k_in = wl.shape
        c4 = self.final_w.shape[0]
        P = k_in // c4
        wl = wl.view(n_out, c4, P).permute(0, 2, 1).reshape(n_out, k_in) * (1.0 / math.sqrt(k_in))
        self.lin_w = wl.contiguous().to(F16)
        self.lin_b = f32('final_linear.bias')
        # decoder (U-Net up path) + SFT heads + plain toRGB
        self.up = []
        for i in range(L):
            p = f'conv_body_up.{i}'
            cs, ch = f'condition_scale.{i}', f'condition_shift.{i}'
            wrgb = g(f'toRGB.{i}.weight').float()
            self.up.append(dict(
                w1=eq(f'{p}.conv1.0.weight'), b1=f32(f'{p}.conv1.1.bias'),
                w2=eq(f'{p}.conv2.weight'), b2=f32(f'{p}.conv2.activation.bias'), ws=eq(f'{p}.skip.weight'),
                w2_fold=ops.upfold_weights(g(f'{p}.conv2.weight').float(),
                                           1.0 / math.sqrt(g(f'{p}.conv2.weight').shape[1] * 9)),
                wh0=torch.cat([eq(f'{cs}.0.weight'), eq(f'{ch}.0.weight')], 0).contiguous(),
                bh0=torch.cat([f32(f'{cs}.0.bias'), f32(f'{ch}.0.bias')]).contiguous(),
                wsc=eq(f'{cs}.2.weight'), bsc=f32(f'{cs}.2.bias'), wsh=eq(f'{ch}.2.weight'), bsh=f32(f'{ch}.2.bias'),
                wrgb=(wrgb[:, :, 0, 0] / math.sqrt(wrgb.shape[1])).contiguous(), brgb=f32(f'toRGB.{i}.bias')))
        # StyleGAN decoder
        pack_decoder(self, net, g, f32)


def pack_decoder(self, net, g, f32, train=False):
    """Packs the StyleGAN2 decoder's weights (stylegan2_ocr_arch.py:408-497) onto `self`: const, sc1, rgb1, sconv[], rgbs[],
    stored_noise[], mod_wscale and the style MLP.  g(key) -> tensor on the device, f32(key) -> contiguous fp32 copy.
    Shared by the inference engine (_Packed) and the training path (train.DecoderState: with fix_decoder=True these never
    change, so they are packed once, not per optimiser step).  train=True adds the adjoint weight packs of the input-gradient
    convs (`w_dgrad`: [cin][9*cout])."""
    dev = self.dev
    L = net.log_size - 2
    self.L = L
    D = 'stylegan_decoder'
    cst = g(f'{D}.constant_input.weight').float()[0]                       # (C, 4, 4r)
    self.const = cst.permute(1, 2, 0).contiguous().to(F16)                 # [4][4r][C]

    def style_conv(p, upsample):
        w = g(f'{p}.modulated_conv.weight').float()[0]                     # (cout, cin, 3, 3)
        cout, cin = w.shape[:2]
        scale = 1.0 / math.sqrt(cin * 9)
        d = dict(cin=cin, cout=cout, scale2=scale * scale, wsq=w.pow(2).sum([2, 3]).contiguous(),
                 mod_w=f32(f'{p}.modulated_conv.modulation.weight'), mod_b=f32(f'{p}.modulated_conv.modulation.bias'),
                 gain=f32(f'{p}.weight'), bias=f32(f'{p}.activate.bias'))
        if upsample:
            ws = (w * scale).to(F16)
            d['w_phase'] = [torch.cat([ws[:, :, kh, kw] for kh, kw in ops.convt_phase_taps(py, px)], 1).contiguous()
                            for py, px in ops.CONVT_PHASES]
            d['w_merged'] = ops.convt_merged_weight(w, scale)
            if train:   # adjoint of conv_transpose2d(stride 2) = the stride-2 conv with in / out channels exchanged (no flip)
                d['w_dgrad'] = _pack_conv(w.permute(1, 0, 2, 3), scale)
        else:
            d['w'] = _pack_conv(w, scale)
            if train:   # adjoint of the stride-1 'same' conv = the same conv with flipped taps and exchanged channels
                d['w_dgrad'] = ops.conv_dgrad_weight(d['w'], cin)
        return d

    def rgb(p):
        w = g(f'{p}.modulated_conv.weight').float()[0, :, :, 0, 0]         # (3, cin)
        return dict(w=(w / math.sqrt(w.shape[1])).contiguous(), bias=f32(f'{p}.bias').reshape(3).contiguous(),
                    mod_w=f32(f'{p}.modulated_conv.modulation.weight'),
                    mod_b=f32(f'{p}.modulated_conv.modulation.bias'))

    self.sc1 = style_conv(f'{D}.style_conv1', False)
    self.rgb1 = rgb(f'{D}.to_rgb1')
    self.sconv = [style_conv(f'{D}.style_convs.{j}', j % 2 == 0) for j in range(2 * L)]
    self.rgbs = [rgb(f'{D}.to_rgbs.{i}') for i in range(L)]
    self.stored_noise = [f32(f'{D}.noises.noise{j}') for j in range(2 * L + 1)]
    self.mod_wscale = 1.0 / math.sqrt(net.num_style_feat)
    # style MLP (only executed when input_is_latent=False)
    self.mlp_w = torch.stack([g(f'{D}.style_mlp.{i}.weight').float() for i in range(1, net.num_mlp + 1)]).contiguous() \
        if net.num_mlp > 0 else torch.zeros(0, net.num_style_feat, net.num_style_feat, device=dev)
    self.mlp_b = torch.stack([g(f'{D}.style_mlp.{i}.bias').float() for i in range(1, net.num_mlp + 1)]).contiguous() \
        if net.num_mlp > 0 else torch.zeros(0, net.num_style_feat, device=dev)
    self.mlp_lr_mul = float(net.stylegan_decoder.style_mlp[1].lr_mul) if net.num_mlp > 0 else 1.0


class PwOp:
    """One prepared launch of a memory-bound kernel: callable + the tensors it reads / writes once (algorithmic HBM
    bytes for the roofline report of bench.py)."""

    def __init__(self, name, fn, *tensors):
        self.name, self.fn = name, fn
        self.nbytes = sum(t.numel() * t.element_size() for t in tensors if t is not None)

    def __call__(self):
        self.fn()


class _Steps(list):
    """Launch list of a plan.  The list itself holds the ops in construction order (callables, or ('rgb', i) markers
    for the optional U-Net toRGB heads); `sched` adds the stream lane of every op and the cross-lane dependencies:
    ('op', lane, op) | ('rec', lane, key) | ('wait', lane, key).  Lane 0 is the caller's stream, lane 1 an auxiliary
    stream: the StyleGAN decoder runs there one level behind the U-Net decoder that feeds it, and the
    1x1 skip branch of every ResBlock runs beside its 3x3 branch."""

    def __init__(self):
        super().__init__()
        self.sched = []
        self.lane = 0

    def append(self, op):
        super().append(op)
        self.sched.append(('op', self.lane, op))

    def rec(self, key, lane=None):
        self.sched.append(('rec', self.lane if lane is None else lane, key))

    def wait(self, key, lane=None):
        self.sched.append(('wait', self.lane if lane is None else lane, key))


class _Plan:
    """All buffers and prepared launches for one batch size."""

    def __init__(self, eng, B):
        net, pk = eng.net, eng.packed
        dev = pk.dev
        self.B = B
        L = pk.L
        H, W = net.input_height, net.input_width
        nf = net.num_style_feat
        steps = _Steps()    # launch list: zero-arg callables (ConvOp / PwOp) or ('rgb', i) markers, with stream lanes
        self.steps = steps
        e16 = lambda *s: torch.empty(*s, device=dev, dtype=F16)  # noqa: E731
        z16 = lambda *s: torch.zeros(*s, device=dev, dtype=F16)  # noqa: E731
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=F32)  # noqa: E731
        inv = ops.INV_SQRT2

        self.x_in = e32(B, 3, H, W)
        c0 = pk.first_w.shape[0]
        feat = e16(B, H, W, c0)
        steps.append(PwOp('first_conv', lambda o=feat: ops.first_conv(self.x_in, pk.first_w, pk.first_b, o), self.x_in, feat))
        # ---------------- encoder: ResBlock x L (stylegan2_ocr_arch.py:729-734)
        skips = []
        h, w = H, W
        for i in range(L):
            d = pk.down[i]
            cin, cout = d['w1'].shape[0], d['w2'].shape[0]
            steps.rec(f'enc_in{i}')
            t1 = e16(B, h, w, cin)
            steps.append(ops.conv_same(feat, d['w1'], t1, 3, bias=d['b1'], act=True))
            p = z16(B, h + 2, w + 2, cin)
            steps.append(PwOp('fir_pad22', lambda a=t1, o=p: ops.fir_pad22(a, o), t1, p))
            steps.lane = 1                     # skip branch (FIR + stride-2 sampling, 1x1 conv) beside the 3x3 branch
            steps.wait(f'enc_in{i}')
            sk_in = e16(B, h // 2, w // 2, cin)
            steps.append(PwOp('fir_down2', lambda a=feat, o=sk_in: ops.fir_down2(a, o), feat, sk_in))
            sk = e16(B, h // 2, w // 2, cout)
            steps.append(ops.conv_same(sk_in, d['ws'], sk, 1))
            steps.rec(f'enc_skip{i}')
            steps.lane = 0
            steps.wait(f'enc_skip{i}')
            nxt = e16(B, h // 2, w // 2, cout)
            oh, ow = h // 2, w // 2
            steps.append(ops.conv3x3_s2(p, h, w, d['w2'], nxt, bias=d['b2'], act=True, res=sk, res_mode=1,
                                        res_strides=(cout, ow * cout, oh * ow * cout), res_wh=(ow, oh), res_scale=inv))
            feat = nxt
            skips.insert(0, feat)
            h, w = oh, ow
        # ---------------- final conv + style code (gfpganv1_ocr_arch.py:358-363)
        c4 = pk.final_w.shape[0]
        g = e16(B, h, w, c4)
        steps.append(ops.conv_same(feat, pk.final_w, g, 3, bias=pk.final_b, act=True))
        n_lin = pk.lin_w.shape[0]
        self.style_code = e32(B, n_lin)
        steps.append(ops.linear_as_conv(g.view(B, -1), pk.lin_w, self.style_code, bias=pk.lin_b,
                                        block_n=64 if n_lin % 64 == 0 else None))
        if not net.input_is_latent:
            # style code -> latent through the style MLP (gfpganv1_ocr_arch.py:77-78, stylegan2_ocr_arch.py:424-430)
            z = self.style_code
            self.style_code = e32(B, n_lin)
            steps.append(lambda a=z, o=self.style_code: ops.style_mlp(a, pk.mlp_w, pk.mlp_b, o, pk.mlp_lr_mul))
        steps.rec('style_code')
        if net.different_w:
            self.num_latent = n_lin // nf
            self.latent = self.style_code.view(B, self.num_latent, nf)
        else:
            self.num_latent = 1
            self.latent = self.style_code.view(B, 1, nf)

        def lat(i):
            return i if net.different_w else 0

        # ---------------- U-Net decoder with SFT heads (gfpganv1_ocr_arch.py:366-378)
        self.cond = []      # (scale, shift) per level, NHWC fp16
        self.out_rgbs = []
        self.rgb_steps = []
        feat = g
        for i in range(L):
            d = pk.up[i]
            cin, cout = d['w1'].shape[0], d['w2'].shape[0]
            a = e16(B, h, w, cin)
            steps.append(PwOp('add', lambda x=feat, y=skips[i], o=a: ops.add(x, y, o), feat, skips[i], a))
            steps.rec(f'up_a{i}')
            # folding pays where the plain conv would run N = cout <= 64 MMAs in the generic kernel (256 -> 64 @64x192:
            # 442 us with the bilinear kernel -> 254 us); wide layers gain nothing from N = 4*cout and the 64 -> 32 layer
            # is faster in the row-sliding kernel (weights resident), see tools/time_ops.py
            fold = (eng.upfold and cin % 64 == 0 and cin >= 128 and cout % 16 == 0 and (cout & (cout - 1)) == 0 and
                    4 * cout <= 256) or (eng.upfold_all and cin % 64 == 0 and (cout & (cout - 1)) == 0 and cout >= 16)
            if fold:
                # ConvUpLayer folded (ops.UpFoldConv): conv1 writes into the interior of a replicate-padded buffer, the
                # up-sampled tensor is never materialised, conv2 runs over the low-resolution input with N = 4*cout
                tp = e16(B, h + 2, w + 2, cin)
                steps.append(ops.ConvOp([ops.nhwc_view(a)], d['w1'], cin, cin, ops.taps_3x3(), (w, h, B), tp[:, 1:, 1:, :],
                                        (cin, (w + 2) * cin, (h + 2) * (w + 2) * cin), bias=d['b1'], act=True))
            else:
                t1 = e16(B, h, w, cin)
                steps.append(ops.conv_same(a, d['w1'], t1, 3, bias=d['b1'], act=True))
                u = e16(B, 2 * h, 2 * w, cin)
                steps.append(PwOp('bilinear_up2', lambda x=t1, o=u: ops.bilinear_up2(x, o), t1, u))
            sl = e16(B, h, w, cout)
            steps.lane = 2
            steps.wait(f'up_a{i}')
            steps.append(ops.conv_same(a, d['ws'], sl, 1))
            steps.rec(f'up_sl{i}')
            steps.lane = 0
            h2, w2 = 2 * h, 2 * w
            feat = e16(B, h2, w2, cout)
            if fold:
                uf = ops.UpFoldConv(tp, d['w2_fold'], d['b2'], feat, sl, inv)
                steps.append(PwOp('replicate_border', uf.pad, tp[:, 0]))
                steps.rec(f'up_tp{i}')
                steps.lane = 2                      # the four border GEMMs + corner fix run beside the main conv's start
                steps.wait(f'up_tp{i}')
                for bop in uf.border_ops:
                    steps.append(bop)
                steps.append(PwOp('upfold_corners', uf.corners))
                steps.rec(f'up_corr{i}')
                steps.lane = 0
                steps.wait(f'up_sl{i}')
                steps.wait(f'up_corr{i}')
                steps.append(uf.main)
            else:
                steps.wait(f'up_sl{i}')
                steps.append(ops.conv_same(u, d['w2'], feat, 3, bias=d['b2'], act=True, res=sl, res_mode=2,
                                           res_strides=(cout, w * cout, h * w * cout), res_wh=(w, h), res_scale=inv))
            c_sft = d['wsc'].shape[0]
            sc, sh = e16(B, h2, w2, c_sft), e16(B, h2, w2, c_sft)
            if cout <= _SFT_SPLIT:
                # 32-channel heads: a merged 64-channel hidden tensor would make each output conv read 64-byte half rows (the
                # whole 128-byte line is fetched: 115 us against 79 us for a contiguous 32-channel input, B = 64 at 128x384),
                # and two 32 -> 32 launches (two CTAs per SM in the row kernel) beat one 32 -> 64 launch (2 x 79 vs 181 us):
                # each head is its own chain of two convs, the shift head on lane 2 beside the scale head.  64-channel heads
                # (64 x 192): four row-kernel launches of ~70 us against 157 us (merged 64 -> 128, generic kernel) + 2 x 72 us,
                # 7.36 -> 7.32 ms per step (three alternating runs each)
                steps.rec(f'up_feat{i}')
                for half, wk, bk, dst in ((0, 'wsc', 'bsc', sc), (1, 'wsh', 'bsh', sh)):
                    if half == 1:
                        steps.lane = 2
                        steps.wait(f'up_feat{i}')
                    hid_h = e16(B, h2, w2, cout)
                    steps.append(ops.conv_same(feat, d['wh0'][half * cout:(half + 1) * cout], hid_h, 3,
                                               bias=d['bh0'][half * cout:(half + 1) * cout], act=True))
                    steps.append(ops.conv_same(hid_h, d[wk], dst, 3, bias=d[bk]))
            else:
                hid = e16(B, h2, w2, 2 * cout)
                steps.append(ops.conv_same(feat, d['wh0'], hid, 3, bias=d['bh0'], act=True))
                steps.rec(f'up_hid{i}')
                for half, wk, bk, dst in ((0, 'wsc', 'bsc', sc), (1, 'wsh', 'bsh', sh)):
                    v = ops.View(hid.data_ptr() + 2 * half * cout, cout, w2, h2, B, 2 * cout, w2 * 2 * cout,
                                 h2 * w2 * 2 * cout)
                    rk = dict(tile=(128, 1, 1), row_mode=1, block_n=c_sft) if ops.row_mode_ok(B, h2, w2, cout, c_sft) else {}
                    if half == 1:                      # the shift head runs beside the scale head
                        steps.lane = 2
                        steps.wait(f'up_hid{i}')
                    steps.append(ops.ConvOp([v], d[wk], cout, c_sft, ops.taps_3x3(), (w2, h2, B), dst,
                                            (c_sft, w2 * c_sft, h2 * w2 * c_sft), bias=d[bk], **rk))
            steps.rec(f'up_sh{i}')
            steps.lane = 0
            steps.wait(f'up_sh{i}')
            self.cond.append((sc, sh))
            steps.rec(f'cond{i}')
            rgb = e32(B, 3, h2, w2)
            self.out_rgbs.append(rgb)
            self.rgb_steps.append(PwOp('to_rgb', lambda x=feat, dd=d, o=rgb: ops.to_rgb(x, dd['wrgb'], None, dd['brgb'], None, o),
                                       feat, rgb))
            steps.append(('rgb', len(self.rgb_steps) - 1))
            h, w = h2, w2
        self._hid_keep = None

        # ---------------- style modulation vectors and demodulation tables (hoisted: depend on the latent only)
        steps.lane = 1
        steps.wait('style_code')
        mod_layers, demod_layers = [], []

        def mod(layer, lat_idx):
            s = e32(B, layer['mod_w'].shape[0])
            mod_layers.append((layer['mod_w'], layer['mod_b'], lat(lat_idx), s))
            return s

        def dem(layer, s):
            dd = e32(B, layer['cout'])
            demod_layers.append((s, layer['wsq'], layer['scale2'], dd))
            return dd

        s_sc1 = mod(pk.sc1, 0)
        d_sc1 = dem(pk.sc1, s_sc1)
        s_rgb1 = mod(pk.rgb1, 1)
        s_conv, d_conv, s_rgb = [], [], []
        for lvl in range(L):
            i = 1 + 2 * lvl
            s1 = mod(pk.sconv[2 * lvl], i)
            s2 = mod(pk.sconv[2 * lvl + 1], i + 1)
            s_conv += [s1, s2]
            d_conv += [dem(pk.sconv[2 * lvl], s1), dem(pk.sconv[2 * lvl + 1], s2)]
            s_rgb.append(mod(pk.rgbs[lvl], i + 2))
        steps.append(ops.ModLinearMulti(self.latent, mod_layers, pk.mod_wscale))
        steps.append(ops.DemodMulti(demod_layers))

        # ---------------- StyleGAN2 decoder with SFT (gfpganv1_ocr_arch.py:108-129)
        self.noise = [None] * (2 * L + 1)   # fp32 [nb,1,h,w] buffers the kernels read; filled per call
        self.noise_shapes = []
        ratio = int(W / H)
        h, w = 4, 4 * ratio
        cch = pk.sc1['cin']
        self.noise_shapes.append((h, w))
        self.noise[0] = e32(B, 1, h, w)
        # per-image ToRGB weights w[o][c] * s_rgb[b][c] for the conv epilogues that fuse ToRGB
        def rgb_weights(layer, s):
            wm = e32(B, 3, layer['w'].shape[1])
            steps.append(PwOp('rgb_wmod', lambda l=layer, ss=s, o=wm: ops.rgb_wmod(l['w'], ss, o), s, wm))
            return wm

        wm_rgb1 = rgb_weights(pk.rgb1, s_rgb1)
        wm_rgbs = [rgb_weights(pk.rgbs[lvl], s_rgb[lvl]) for lvl in range(L)]

        xs = e16(B, h, w, cch)
        steps.append(PwOp('modulate_const', lambda o=xs: ops.modulate_const(pk.const, s_sc1, o), xs))
        # style_conv1 + to_rgb1 (gfpganv1_ocr_arch.py:108-110): the conv epilogue accumulates the ToRGB dot products
        # and writes its output already multiplied by the modulation of the next conv
        last = L == 0
        xs_next = None if last else e16(B, h, w, pk.sc1['cout'])
        op = ops.conv_same(xs, pk.sc1['w'], xs_next, 3, bias=pk.sc1['bias'], demod=d_sc1, noise=self.noise[0],
                           noise_gain=pk.sc1['gain'], noise_strides=(h * w, w), act=True,
                           out_scale=None if last else s_conv[0])
        part = op.attach_rgb(wm_rgb1, (h, w), no_store=last)
        steps.append(op)
        skip = e32(B, 3, h, w)
        steps.append(PwOp('rgb_combine', lambda pt=part, o=skip: ops.rgb_combine(pt, pk.rgb1['bias'], None, o), part, skip))
        xs = xs_next
        for lvl in range(L):
            c1, c2 = pk.sconv[2 * lvl], pk.sconv[2 * lvl + 1]
            cout = c1['cout']
            h2, w2 = 2 * h, 2 * w
            raw = z16(B, h2 + 2, w2 + 2, cout)
            # merged form: N = 256 MMAs for the 64/128-channel levels and one launch for the small ones; the 512-channel
            # levels with enough tiles run faster as four exact-work phase GEMMs (tools/time_convt.py)
            if eng.convt_merged and (cout <= eng.convt_merged_maxc or B * h * w < 40000 or eng.convt_merged_all):
                # one implicit GEMM for the whole stride-2 transposed conv (phases = column blocks)
                steps.append(ops.convt_s2_merged(xs, c1['w_merged'], raw, d_conv[2 * lvl]))
            else:
                steps.rec(f'sg_xs{lvl}')
                for pi, (py, px) in enumerate(ops.CONVT_PHASES):
                    if pi == 1:                    # the three smaller output phases run beside the 4-tap phase
                        steps.lane = 3
                        steps.wait(f'sg_xs{lvl}')
                    steps.append(ops.convt_s2_phase(xs, c1['w_phase'][pi], py, px, raw, demod=d_conv[2 * lvl]))
                steps.rec(f'sg_ph{lvl}')
                steps.lane = 1
                steps.wait(f'sg_ph{lvl}')
            n1 = e32(B, 1, h2, w2)
            n2 = e32(B, 1, h2, w2)
            self.noise[2 * lvl + 1], self.noise[2 * lvl + 2] = n1, n2
            self.noise_shapes += [(h2, w2), (h2, w2)]
            sc, sh = self.cond[lvl]
            steps.wait(f'cond{lvl}')
            xs2 = e16(B, h2, w2, cout)
            steps.append(PwOp('upfir_act', lambda r=raw, o=xs2, n=n1, c=c1, a=sc, b_=sh, sn=s_conv[2 * lvl + 1]:
                              ops.upfir_act(r, o, n, o.shape[1] * o.shape[2], c['gain'], c['bias'], a, b_, a.shape[3], sn),
                              raw, xs2, n1, sc, sh))
            last = lvl == L - 1
            xs = None if last else e16(B, h2, w2, cout)
            op = ops.conv_same(xs2, c2['w'], xs, 3, bias=c2['bias'], demod=d_conv[2 * lvl + 1], noise=n2,
                               noise_gain=c2['gain'], noise_strides=(h2 * w2, w2), act=True,
                               out_scale=None if last else s_conv[2 * lvl + 2])
            part = op.attach_rgb(wm_rgbs[lvl], (h2, w2), no_store=last)
            steps.append(op)
            nskip = e32(B, 3, h2, w2)
            steps.append(PwOp('rgb_combine', lambda pt=part, r=pk.rgbs[lvl], sk=skip, o=nskip: ops.rgb_combine(pt, r['bias'], sk, o),
                              part, skip, nskip))
            skip = nskip
            h, w = h2, w2
        steps.rec('image')
        steps.lane = 0
        steps.wait('image')
        self.image = skip
        self.graphs = {}
        import os
        self.lanes = max(1, min(4, int(os.environ.get('B200IR_LANES', '4'))))
        self.aux_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        self.events = {e[2]: torch.cuda.Event() for e in steps.sched if e[0] == 'rec'}
        self.noise_state = None

    def launch(self, return_rgb):
        """Enqueues one forward on the current stream (lane 0) and the plan's auxiliary stream (lane 1); lane 1 forks
        from the current stream at entry and is joined before return, so callers (and CUDA graph capture) see a
        single-stream operation."""
        cur = torch.cuda.current_stream()
        # lane -> stream; lanes beyond the configured count fold onto lower ones (lane 1 <- 3, lane 0 <- 2)
        n = self.lanes
        lanes = [cur] + self.aux_streams
        if n < 4:
            lanes[3] = lanes[1]
        if n < 3:
            lanes[2] = lanes[0]
        if n < 2:
            lanes[1] = lanes[3] = lanes[0]
        used = [st for st in dict.fromkeys(lanes[1:]) if st is not cur]
        for st in used:
            st.wait_stream(cur)
        for e in self.steps.sched:
            kind, stream = e[0], lanes[e[1]]
            if kind == 'op':
                st = e[2]
                if isinstance(st, tuple):
                    if not return_rgb:
                        continue
                    st = self.rgb_steps[st[1]]
                if stream is cur:
                    st()
                else:
                    with torch.cuda.stream(stream):
                        st()
            elif kind == 'rec':
                self.events[e[2]].record(stream)
            else:
                stream.wait_event(self.events[e[2]])
        for st in used:
            cur.wait_stream(st)


class OcrEngine:
    def __init__(self, net):
        if not net.input_is_latent and net.different_w:
            # the reference fails here too: style_mlp's fused_leaky_relu broadcasts its bias over dim 1 of a
            # (B, num_latent, F) tensor (stylegan2_ocr_arch.py:175, fused_act.py:94)
            raise ValueError('input_is_latent=False needs different_w=False (one style vector per image)')
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('image_restoration_b200 needs the module on a CUDA B200 (no CPU path)')
        from . import _lib
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200ir_device_check(), 'device check')
        self.net = net
        self._sig = self._signature()
        with torch.cuda.device(dev), torch.no_grad():
            self.packed = _Packed(net)
        self.plans = {}
        self.use_graphs = True
        import os
        self.convt_merged = os.environ.get('B200IR_CONVT_MERGED', '1') != '0'
        self.convt_merged_all = os.environ.get('B200IR_CONVT_MERGED', '1') == '2'   # experiment: merged form at every level
        self.convt_merged_maxc = int(os.environ.get('B200IR_CONVT_MERGED_MAXC', '128'))   # large levels: merged up to this cout
        self.upfold = os.environ.get('B200IR_UPFOLD', '1') != '0'
        self.upfold_all = os.environ.get('B200IR_UPFOLD', '1') == '2'     # experiment: fold every ConvUpLayer

    def _signature(self):
        ps = list(self.net.parameters()) + list(self.net.buffers())
        return (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))

    def stale(self):
        return self._sig != self._signature()

    MAX_PLANS = 4   # a plan holds every activation of its batch size (~0.15 GB per crop @128x384): keep the most recent

    def plan(self, B):
        if B in self.plans:
            self.plans[B] = self.plans.pop(B)          # most recently used last
        else:
            while len(self.plans) >= self.MAX_PLANS:
                self.plans.pop(next(iter(self.plans)))  # evict the least recently used plan (buffers, graphs)
            self.plans[B] = _Plan(self, B)
        return self.plans[B]

    @torch.no_grad()
    def forward(self, x, return_rgb=True, randomize_noise=True, save_feat_path=None, load_feat_path=None,
                noise=None, uint8_io=False, bgr=True):
        """uint8_io: `x` is a uint8 HWC image batch (B,H,W,3) (BGR when `bgr`) as the serving scripts hold it; the
        img2tensor/normalize on the way in and tensor2img on the way out (api.py:96-105) run on the device and the
        returned image is uint8 (B,H,W,3)."""
        net = self.net
        B = x.shape[0]
        if uint8_io:
            if x.dtype != torch.uint8 or tuple(x.shape[1:]) != (net.input_height, net.input_width, 3):
                raise ValueError(f'uint8_io expects uint8 (B,{net.input_height},{net.input_width},3), got '
                                 f'{x.dtype} {tuple(x.shape)}')
        elif tuple(x.shape[1:]) != (3, net.input_height, net.input_width):
            raise ValueError(f'expected input (B,3,{net.input_height},{net.input_width}), got {tuple(x.shape)}')
        dev = self.packed.dev
        if B == 0:                                       # empty batch: nothing to launch (as nn.Module layers return empties)
            H, W = net.input_height, net.input_width
            if uint8_io:
                return torch.empty(0, H, W, 3, device=dev, dtype=torch.uint8), []
            L = self.packed.L
            rgbs = [torch.empty(0, 3, 8 * 2 ** i, 8 * 2 ** i * int(W / H), device=dev, dtype=x.dtype) for i in range(L)] \
                if return_rgb else []
            return torch.empty(0, 3, H, W, device=dev, dtype=x.dtype), rgbs
        with torch.cuda.device(dev):
            plan = self.plan(B)
            if uint8_io:
                ops.u8_to_input(x.contiguous(), plan.x_in, swap_rb=bgr)
            else:
                plan.x_in.copy_(x)
            pk = self.packed
            if noise is None and not randomize_noise and plan.noise_state == 'stored':
                pass                                  # buffers already hold the registered noise planes
            else:
                for j, buf in enumerate(plan.noise):
                    if noise is not None:
                        buf.copy_(noise[j].expand_as(buf))
                    elif randomize_noise:
                        buf.normal_()
                    else:
                        buf.copy_(pk.stored_noise[j].expand_as(buf))
                plan.noise_state = 'stored' if (noise is None and not randomize_noise) else 'other'
            if load_feat_path is not None or save_feat_path is not None or not self.use_graphs:
                self._run_eager(plan, return_rgb, save_feat_path, load_feat_path)
            else:
                key = bool(return_rgb)
                gr = plan.graphs.get(key)
                if gr is None:
                    plan.launch(return_rgb)          # warm-up (also surfaces launch errors outside capture)
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        plan.launch(return_rgb)
                    plan.graphs[key] = gr
                gr.replay()
            rgbs = [t.clone() for t in plan.out_rgbs] if return_rgb else []
            if uint8_io:
                image = torch.empty(B, net.input_height, net.input_width, 3, device=dev, dtype=torch.uint8)
                ops.image_to_u8(plan.image, image, swap_rb=bgr)
                return image, rgbs
            image = plan.image.clone()
        return image.to(x.dtype) if x.dtype != F32 else image, rgbs

    def _run_eager(self, plan, return_rgb, save_feat_path, load_feat_path):
        """Eager (non-graph) execution; also serves save_feat_path / load_feat_path (gfpganv1_ocr_arch.py:380-384):
        the SFT conditions are saved as the reference saves them (list of 2L fp32 NCHW tensors, scale then shift per
        level) and loaded conditions replace the computed ones before the StyleGAN decoder consumes them."""
        plan.launch(return_rgb)
        if save_feat_path is not None:
            conds = []
            for sc, sh in plan.cond:
                for t in (sc, sh):
                    b, h, w, c = t.shape
                    o = torch.empty(b, c, h, w, device=t.device, dtype=F32)
                    ops.nhwc_to_nchw_f32(t, o)
                    conds.append(o)
            torch.save(conds, save_feat_path)
        if load_feat_path is not None:
            conds = torch.load(load_feat_path)
            if len(conds) != 2 * len(plan.cond):
                raise ValueError(f'{load_feat_path}: expected {2 * len(plan.cond)} condition tensors, got {len(conds)}')
            for i, (sc, sh) in enumerate(plan.cond):
                for t, src in ((sc, conds[2 * i]), (sh, conds[2 * i + 1])):
                    src = src.to(t.device, F32).contiguous()
                    if tuple(src.shape) != (t.shape[0], t.shape[3], t.shape[1], t.shape[2]):
                        raise ValueError(f'{load_feat_path}: condition {i} has shape {tuple(src.shape)}')
                    ops.nchw_to_nhwc_pad(src, t)
            # re-run the StyleGAN decoder (lanes 1 and 3 of the schedule) on the loaded conditions
            for e in plan.steps.sched:
                if e[0] == 'op' and e[1] in (1, 3):
                    e[2]()
