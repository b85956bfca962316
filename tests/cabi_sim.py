"""TEST INFRASTRUCTURE (not product code): a CPU functional simulator of the libb200ir C ABI (include/b200ir.h).

The product has no CPU path: `_lib.lib()` loads libb200ir.so and every entry point needs an sm_100 GPU.  The HOST logic above
the ABI, however — the launch order of the training step, which buffer feeds which kernel, the adjoint chain of the frozen
StyleGAN2 decoder, weight packing, loss scaling, the optimiser wiring — is Python, and its correctness does not depend on the
GPU.  `install()` swaps the ctypes handle for this object, whose methods restate each entry point's documented semantics with
torch CPU ops on the raw pointers they are handed (CPU tensors' data_ptr() are host addresses), so that host logic can be
checked against the fp32 oracle's autograd in the `-m "not gpu"` suite.  The kernels themselves are checked on the GPU
(`-m gpu`), each against a torch restatement; nothing here is imported by the package.

Only the entry points the training step uses are simulated; anything else raises.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)
_DT = {torch.float16: 2, torch.float32: 4, torch.uint8: 1, torch.int32: 4, torch.float64: 8}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, C.c_void_p):
        return p.value or 0
    if hasattr(p, '_obj'):                 # C.byref(x)
        return C.addressof(p._obj)
    raise TypeError(type(p))


def _obj(p):
    return p._obj if hasattr(p, '_obj') else p


def T(p, shape, dtype):
    """Contiguous tensor view of raw host memory (shares storage: writes land in the caller's buffer)."""
    a = _addr(p)
    if a == 0:
        return None
    n = 1
    for s in shape:
        n *= int(s)
    buf = (C.c_char * (n * _DT[dtype])).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=n).view(*[int(s) for s in shape])


def TS(p, shape, strides, dtype):
    """Strided view (element strides) of raw host memory."""
    a = _addr(p)
    span = 1 + sum((int(s) - 1) * int(st) for s, st in zip(shape, strides))
    buf = (C.c_char * (span * _DT[dtype])).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=span).as_strided([int(s) for s in shape], [int(s) for s in strides])


def fir2d():
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None] * k[:, None]
    return k / k.sum()


def upfirdn(x, kernel, up=1, down=1, pad=(0, 0)):
    """NCHW upfirdn2d (upfirdn2d.py:162-192)."""
    b, c, h, w = x.shape
    if up > 1:
        z = x.new_zeros(b, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.view(b, c, h * up, w * up)
    x = F.pad(x, [pad[0], pad[1], pad[0], pad[1]])
    kf = torch.flip(kernel, [0, 1]).to(x)[None, None]
    y = F.conv2d(x.reshape(b * c, 1, x.shape[2], x.shape[3]), kf)
    return y.view(b, c, y.shape[2], y.shape[3])[:, :, ::down, ::down]


def nhwc(t):
    return t.permute(0, 2, 3, 1)


def nchw(t):
    return t.permute(0, 3, 1, 2)


class SimLib:
    def __init__(self):
        self.launches = 0
        self.err = b''
        self.plans = {}
        self.next_plan = 1

    # ------------------------------------------------------------------ glue
    def b200ir_last_error(self):
        return self.err

    def b200ir_abi_version(self):
        return 1

    def b200ir_launch_count(self):
        return self.launches

    def b200ir_device_check(self):
        return 0

    def __getattr__(self, name):
        raise AttributeError(f'cabi_sim: {name} is not simulated')

    # ------------------------------------------------------------------ conv
    def b200ir_conv_plan_create(self, d, plan):
        import copy
        desc = _obj(d)
        snap = type(desc).from_buffer_copy(bytes(desc))
        h = self.next_plan
        self.next_plan += 1
        self.plans[h] = snap
        _obj(plan).value = h
        return 0

    def b200ir_conv_plan_launch(self, plan, stream):
        h = plan.value if isinstance(plan, C.c_void_p) else plan
        return self.b200ir_conv_igemm(self.plans[h], stream)

    def b200ir_conv_plan_destroy(self, plan):
        self.plans.pop(plan.value if isinstance(plan, C.c_void_p) else plan, None)

    def b200ir_conv_igemm(self, d, stream):
        d = _obj(d)
        self.launches += 1
        cin, cout, nt = d.cin, d.cout, d.num_taps
        mb, mh, mw = d.m_b, d.m_h, d.m_w
        W = T(d.weight, (mb if d.w_per_image else 1, cout, nt * cin), torch.float16).float()
        views = []
        for v in range(d.num_views):
            a = d.a[v]
            views.append(TS(a.ptr, (a.b, a.h, a.w, a.c), (a.stride_b, a.stride_h, a.stride_w, 1), torch.float16))
        acc = torch.zeros(mb, mh, mw, cout)
        n_tiles = cout // d.block_n
        for t in range(nt):
            v = views[d.tap_view[t]]
            vb, vh, vw, _ = v.shape
            dx, dy = d.tap_dx[t], d.tap_dy[t]
            A = torch.zeros(mb, mh, mw, cin)
            y0, y1 = max(0, -dy), min(mh, vh - dy)
            x0, x1 = max(0, -dx), min(mw, vw - dx)
            b1 = min(mb, vb)
            if y1 > y0 and x1 > x0:
                A[:b1, y0:y1, x0:x1] = v[:b1, y0 + dy:y1 + dy, x0 + dx:x1 + dx, :cin].float()
            Wt = W[:, :, t * cin:(t + 1) * cin]                                   # [mb | 1, cout, cin]
            part = torch.einsum('bhwi,boi->bhwo', A, Wt.expand(mb, cout, cin))
            if d.use_tap_mask:
                for j in range(n_tiles):
                    if not (d.tap_mask[j] >> t) & 1:
                        part[..., j * d.block_n:(j + 1) * d.block_n] = 0
            acc += part
        n_idx = torch.arange(cout)
        ps_c = d.ps_c if d.ps_c else d.block_n
        demod_c = d.demod_c if d.demod_c else cout
        v = acc
        if d.demod:
            dm = T(d.demod, (mb, demod_c), torch.float32)
            v = v * dm[:, n_idx % demod_c].view(mb, 1, 1, cout)
        if d.corr_top or d.rgb_w:
            raise NotImplementedError('cabi_sim: folded ConvUpLayer / fused ToRGB epilogues are not simulated')
        ys = torch.arange(mh) * (d.ps_r if d.ps_r else d.out_y_mul) + (0 if d.ps_r else d.out_y_off)
        xs = torch.arange(mw) * (d.ps_r if d.ps_r else d.out_x_mul) + (0 if d.ps_r else d.out_x_off)
        if d.noise:
            assert not d.ps_r
            gain = T(d.noise_gain, (1,), torch.float32)[0]
            nzp = TS(d.noise, (mb, mh, mw), (d.noise_stride_b, d.noise_stride_y * d.out_y_mul, d.out_x_mul), torch.float32) \
                if (d.out_y_off == 0 and d.out_x_off == 0) else None
            assert nzp is not None
            v = v + gain * nzp.unsqueeze(-1)
        if d.bias:
            v = v + T(d.bias, (cout,), torch.float32)
        if d.act == 1:
            v = F.leaky_relu(v, 0.2) * SQRT2
        elif d.act == 2:
            v = torch.maximum(v, d.act_slope * v)
        if d.res_mode:
            assert not d.ps_r
            res_mul = d.res_mul if d.res_mul != 0.0 else d.res_scale
            if d.res_mode == 1:
                r = TS(d.res, (mb, mh, mw, cout), (d.res_stride_b, d.res_stride_y * d.out_y_mul, d.res_stride_x * d.out_x_mul, 1),
                       torch.float16).float()
            else:
                lo = TS(d.res, (mb, d.res_h, d.res_w, cout), (d.res_stride_b, d.res_stride_y, d.res_stride_x, 1),
                        torch.float16).float()
                r = nhwc(F.interpolate(nchw(lo), scale_factor=2, mode='bilinear', align_corners=False))
                assert r.shape[1] == mh and r.shape[2] == mw
            v = v * d.res_scale + r * res_mul
        if d.out_scale:
            v = v * T(d.out_scale, (mb, cout), torch.float32).view(mb, 1, 1, cout)
        if d.no_store:
            return 0
        odt = torch.float32 if d.out_fp32 else torch.float16
        if d.ps_r:
            r = d.ps_r
            for t in range(cout // ps_c):
                ty, tx = t // r, t % r
                o = TS(_addr(d.out) + _DT[odt] * (ty * d.out_stride_y + tx * d.out_stride_x + d.out_c_off),
                       (mb, mh, mw, ps_c), (d.out_stride_b, d.out_stride_y * r, d.out_stride_x * r, 1), odt)
                o.copy_(v[..., t * ps_c:(t + 1) * ps_c].clamp(-65504, 65504))
        else:
            o = TS(_addr(d.out) + _DT[odt] * (d.out_y_off * d.out_stride_y + d.out_x_off * d.out_stride_x + d.out_c_off),
                   (mb, mh, mw, cout), (d.out_stride_b, d.out_stride_y * d.out_y_mul, d.out_stride_x * d.out_x_mul, 1), odt)
            o.copy_(v.clamp(-65504, 65504) if odt == torch.float16 else v)
        return 0

    # ------------------------------------------------------------------ forward memory-bound stages
    def b200ir_first_conv(self, x, w, bias, out, B, H, W, cout, stream):
        self.launches += 1
        xt = T(x, (B, 3, H, W), torch.float32)
        wt = T(w, (cout, 3), torch.float32)
        y = torch.einsum('bkhw,ck->bhwc', xt, wt) + T(bias, (cout,), torch.float32)
        T(out, (B, H, W, cout), torch.float16).copy_(F.leaky_relu(y, 0.2) * SQRT2)
        return 0

    def b200ir_fir_pad22(self, inp, out, B, H, W, Cc, out_h, out_w, stream):
        self.launches += 1
        x = T(inp, (B, H, W, Cc), torch.float16).float()
        y = nhwc(upfirdn(nchw(x), fir2d(), pad=(2, 2)))
        T(out, (B, out_h, out_w, Cc), torch.float16)[:, :H + 1, :W + 1].copy_(y)
        return 0

    def b200ir_fir_pad11(self, inp, out, B, H, W, Cc, in_h, in_w, stream):
        self.launches += 1
        x = T(inp, (B, in_h, in_w, Cc), torch.float16)[:, :H + 1, :W + 1].float()
        T(out, (B, H, W, Cc), torch.float16).copy_(nhwc(upfirdn(nchw(x), fir2d(), pad=(1, 1))))
        return 0

    def b200ir_fir_down2(self, inp, out, B, H, W, Cc, stream):
        self.launches += 1
        x = T(inp, (B, H, W, Cc), torch.float16).float()
        T(out, (B, H // 2, W // 2, Cc), torch.float16).copy_(nhwc(upfirdn(nchw(x), fir2d(), down=2, pad=(1, 1))))
        return 0

    def b200ir_fir_down2_adjoint(self, d, add, out, B, h, w, Cc, stream):
        self.launches += 1
        x = T(d, (B, h, w, Cc), torch.float16).float()
        y = nhwc(upfirdn(nchw(x), fir2d(), up=2, pad=(2, 1)))
        if _addr(add):
            y = y + T(add, (B, 2 * h, 2 * w, Cc), torch.float16).float()
        T(out, (B, 2 * h, 2 * w, Cc), torch.float16).copy_(y)
        return 0

    def b200ir_bilinear_up2(self, inp, out, B, h, w, Cc, stream):
        self.launches += 1
        x = T(inp, (B, h, w, Cc), torch.float16).float()
        T(out, (B, 2 * h, 2 * w, Cc), torch.float16).copy_(
            nhwc(F.interpolate(nchw(x), scale_factor=2, mode='bilinear', align_corners=False)))
        return 0

    def b200ir_bilinear_up2_adjoint(self, d, out, B, h, w, Cc, scale, stream):
        self.launches += 1
        g = T(d, (B, 2 * h, 2 * w, Cc), torch.float16).float()
        with torch.enable_grad():      # simulated kernels run inside autograd.Function.backward, where grad mode is off
            lo = torch.zeros(B, Cc, h, w, requires_grad=True)
            F.interpolate(lo, scale_factor=2, mode='bilinear', align_corners=False).backward(nchw(g))
        T(out, (B, h, w, Cc), torch.float16).copy_(nhwc(lo.grad) * scale)
        return 0

    def b200ir_add(self, a, b, out, n, stream):
        self.launches += 1
        T(out, (n,), torch.float16).copy_(T(a, (n,), torch.float16).float() + T(b, (n,), torch.float16).float())
        return 0

    def b200ir_upfir_act(self, raw, out, B, h2, w2, Cc, raw_h, raw_w, noise, noise_sb, noise_gain, bias, scale, shift, c_sft,
                         s_next, stream):
        self.launches += 1
        r = T(raw, (B, raw_h, raw_w, Cc), torch.float16)[:, :h2 + 1, :w2 + 1].float()
        y = nhwc(upfirdn(nchw(r), fir2d() * 4, pad=(1, 1)))
        if _addr(noise):
            y = y + T(noise_gain, (1,), torch.float32)[0] * TS(noise, (B, h2, w2), (noise_sb, w2, 1), torch.float32).unsqueeze(-1)
        y = F.leaky_relu(y + T(bias, (Cc,), torch.float32), 0.2) * SQRT2
        if _addr(scale):
            sc = T(scale, (B, h2, w2, c_sft), torch.float16).float()
            sh = T(shift, (B, h2, w2, c_sft), torch.float16).float()
            y = torch.cat([y[..., :Cc - c_sft], y[..., Cc - c_sft:] * sc + sh], -1)
        if _addr(s_next):
            y = y * T(s_next, (B, Cc), torch.float32).view(B, 1, 1, Cc)
        T(out, (B, h2, w2, Cc), torch.float16).copy_(y.clamp(-65504, 65504))
        return 0

    def b200ir_to_rgb(self, x, B, h, w, Cc, wrgb, s, bias, skip, rgb, s_next, xs_out, stream):
        self.launches += 1
        xt = T(x, (B, h, w, Cc), torch.float16).float()
        wm = T(wrgb, (3, Cc), torch.float32).unsqueeze(0)
        if _addr(s):
            wm = wm * T(s, (B, Cc), torch.float32).unsqueeze(1)
        o = torch.einsum('bhwc,boc->bohw', xt, wm.expand(B, 3, Cc)) + T(bias, (3,), torch.float32).view(1, 3, 1, 1)
        if _addr(skip):
            o = o + upfirdn(T(skip, (B, 3, h // 2, w // 2), torch.float32), fir2d() * 4, up=2, pad=(2, 1))
        T(rgb, (B, 3, h, w), torch.float32).copy_(o)
        if _addr(xs_out):
            T(xs_out, (B, h, w, Cc), torch.float16).copy_(xt * T(s_next, (B, Cc), torch.float32).view(B, 1, 1, Cc))
        return 0

    def b200ir_modulate_const(self, cst, s, out, B, P, Cc, stream):
        self.launches += 1
        T(out, (B, P, Cc), torch.float16).copy_(T(cst, (P, Cc), torch.float16).float().unsqueeze(0) *
                                                 T(s, (B, Cc), torch.float32).unsqueeze(1))
        return 0

    def b200ir_mod_linear_multi(self, latent, L, Fd, layers, n_layers, max_cin, wscale, B, stream):
        from image_restoration_b200._lib import ModLayer
        self.launches += 1
        lat = T(latent, (B, L, Fd), torch.float32)
        recs = (ModLayer * n_layers).from_address(_addr(layers))
        for r in recs:
            w = T(r.w, (r.cin, Fd), torch.float32)
            T(r.s, (B, r.cin), torch.float32).copy_(lat[:, r.lat_idx] @ w.t() * wscale + T(r.bias, (r.cin,), torch.float32))
        return 0

    def b200ir_demod_multi(self, layers, n_layers, max_cout, B, stream):
        from image_restoration_b200._lib import DemodLayer
        self.launches += 1
        recs = (DemodLayer * n_layers).from_address(_addr(layers))
        for r in recs:
            s = T(r.s, (B, r.cin), torch.float32)
            wsq = T(r.wsq, (r.cout, r.cin), torch.float32)
            T(r.d, (B, r.cout), torch.float32).copy_(torch.rsqrt(r.scale2 * (s * s) @ wsq.t() + 1e-8))
        return 0

    def b200ir_minibatch_stddev(self, x, s, out, B, P, Cc, c_pad, group, stream):
        self.launches += 1
        xt = T(x, (B, P, Cc), torch.float16).float()
        M = B // group
        g = xt.view(group, M, P, Cc)
        sd = torch.sqrt(g.var(0, unbiased=False) + 1e-8).mean(dim=(1, 2))            # [M]
        T(s, (M,), torch.float32).copy_(sd)
        o = T(out, (B, P, c_pad), torch.float16)
        o.zero_()
        o[..., :Cc].copy_(xt)
        o[..., Cc].copy_(sd.repeat(group).view(B, 1).expand(B, P))
        return 0

    def b200ir_minibatch_stddev_bwd(self, x, dcat, ds, dx, B, P, Cc, c_pad, group, stream):
        self.launches += 1
        M = B // group
        with torch.enable_grad():
            xt = T(x, (B, P, Cc), torch.float16).float().requires_grad_()
            g = xt.view(group, M, P, Cc)
            sd = torch.sqrt(g.var(0, unbiased=False) + 1e-8).mean(dim=(1, 2))
            (sd * T(ds, (M,), torch.float32)).sum().backward()
        T(dx, (B, P, Cc), torch.float16).copy_(xt.grad + T(dcat, (B, P, c_pad), torch.float16)[..., :Cc].float())
        return 0

    # ------------------------------------------------------------------ backward kernels of round 1
    def b200ir_lrelu_bias_bwd(self, dy, y, dz, dbias, n_pix, Cc, slope, scale, stream):
        self.launches += 1
        d = T(dy, (n_pix, Cc), torch.float16).float()
        if _addr(y):
            yt = T(y, (n_pix, Cc), torch.float16).float()
            v = d * torch.where(yt > 0, scale, scale * slope)
        else:
            v = d * scale
        if _addr(dbias):
            T(dbias, (Cc,), torch.float32).copy_(v.sum(0))
        if _addr(dz):
            T(dz, (n_pix, Cc), torch.float16).copy_(v)
        return 0

    def b200ir_conv_wgrad_view(self, xv, dy, dw, B, H, W, cout, tap_mask, stream):
        self.launches += 1
        a = _obj(xv)
        cin = a.c
        v = TS(a.ptr, (a.b, a.h, a.w, a.c), (a.stride_b, a.stride_h, a.stride_w, 1), torch.float16).float()
        d = T(dy, (B, H, W, cout), torch.float16).float()
        out = T(dw, (cout, 9, cin), torch.float32)
        out.zero_()
        for kh in range(3):
            for kw in range(3):
                if not (tap_mask >> (kh * 3 + kw)) & 1:
                    continue
                dyo, dxo = kh - 1, kw - 1
                A = torch.zeros(B, H, W, cin)
                y0, y1 = max(0, -dyo), min(H, a.h - dyo)
                x0, x1 = max(0, -dxo), min(W, a.w - dxo)
                if y1 > y0 and x1 > x0:
                    A[:, y0:y1, x0:x1] = v[:B, y0 + dyo:y1 + dyo, x0 + dxo:x1 + dxo]
                out[:, kh * 3 + kw] = torch.einsum('bhwo,bhwi->oi', d, A)
        return 0

    def b200ir_conv_wgrad(self, x, dy, dw, B, H, W, cin, cout, stream):
        from image_restoration_b200._lib import View
        v = View(_addr(x), cin, W, H, B, cin, W * cin, H * W * cin)
        return self.b200ir_conv_wgrad_view(v, dy, dw, B, H, W, cout, 0x1FF, stream)

    def b200ir_wgrad_unfold(self, G, dw, f, cin, cout, stream):
        self.launches += 1
        g = T(G, (f, cout, 3, 3, f, cin), torch.float32)            # [s_o, co, kh, kw', s_i, ci]
        out = T(dw, (cout, 3, 3, cin), torch.float32)
        out.zero_()
        for d in (-1, 0, 1):
            for so in range(f):
                t = so + d
                dp, si = (0, t) if 0 <= t < f else ((1, 0) if t == f else (-1, f - 1))
                out[:, :, d + 1] += g[so, :, :, dp + 1, si]
        return 0

    def b200ir_first_conv_wgrad(self, x, dz, dw, B, H, W, cout, stream):
        self.launches += 1
        T(dw, (cout, 3), torch.float32).copy_(torch.einsum('bhwc,bkhw->ck', T(dz, (B, H, W, cout), torch.float16).float(),
                                                           T(x, (B, 3, H, W), torch.float32)))
        return 0

    def b200ir_adam_step(self, param, grad, m, v, n, lr, b1, b2, eps, wd, step, grad_scale, ema, ema_decay, stream):
        self.launches += 1
        p, g = T(param, (n,), torch.float32), T(grad, (n,), torch.float32) * grad_scale
        mt, vt = T(m, (n,), torch.float32), T(v, (n,), torch.float32)
        g = g + wd * p
        g = torch.where(torch.isfinite(g), g, torch.zeros_like(g))
        mt.mul_(b1).add_(g, alpha=1 - b1)
        vt.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.sub_(lr / (1 - b1 ** step) * mt / (vt.sqrt() / math.sqrt(1 - b2 ** step) + eps))
        if _addr(ema):
            e = T(ema, (n,), torch.float32)
            e.mul_(ema_decay).add_(p, alpha=1 - ema_decay)
        return 0

    # ------------------------------------------------------------------ training step kernels (train_ops.cu)
    def b200ir_sft_mod(self, a, scale, shift, c_sft, s_next, out, B, P, Cc, stream):
        self.launches += 1
        v = T(a, (B, P, Cc), torch.float16).float()
        if _addr(scale):
            sc, sh = T(scale, (B, P, c_sft), torch.float16).float(), T(shift, (B, P, c_sft), torch.float16).float()
            v = torch.cat([v[..., :Cc - c_sft], v[..., Cc - c_sft:] * sc + sh], -1)
        if _addr(s_next):
            v = v * T(s_next, (B, Cc), torch.float32).unsqueeze(1)
        T(out, (B, P, Cc), torch.float16).copy_(v.clamp(-65504, 65504))
        return 0

    def b200ir_sft_mod_bwd(self, g, a, a_stride_b, scale, shift, c_sft, s_next, da, accumulate, dscale, dshift, ds, B, P, Cc,
                           stream):
        self.launches += 1
        gv = T(g, (B, P, Cc), torch.float16).float()
        av = TS(a, (B, P, Cc), (a_stride_b, Cc, 1), torch.float16).float()
        o = av
        if _addr(scale):
            sc, sh = T(scale, (B, P, c_sft), torch.float16).float(), T(shift, (B, P, c_sft), torch.float16).float()
            o = torch.cat([av[..., :Cc - c_sft], av[..., Cc - c_sft:] * sc + sh], -1)
        if _addr(ds):
            T(ds, (B, Cc), torch.float32).add_((gv * o).sum(1))
        do = gv * (T(s_next, (B, Cc), torch.float32).unsqueeze(1) if _addr(s_next) else 1.0)
        dav = do
        if _addr(scale):
            T(dscale, (B, P, c_sft), torch.float16).copy_(do[..., Cc - c_sft:] * av[..., Cc - c_sft:])
            T(dshift, (B, P, c_sft), torch.float16).copy_(do[..., Cc - c_sft:])
            dav = torch.cat([do[..., :Cc - c_sft], do[..., Cc - c_sft:] * sc], -1)
        if _addr(da):
            dt = T(da, (B, P, Cc), torch.float16)
            dt.copy_((dt.float() + dav) if accumulate else dav)
        return 0

    def b200ir_style_act_bwd(self, da, a, noise, noise_sb, noise_gain, bias, oscale, mul, out, dd, B, P, Cc, stream):
        self.launches += 1
        dv, av = T(da, (B, P, Cc), torch.float16).float(), T(a, (B, P, Cc), torch.float16).float()
        pos = av > 0
        dz = dv * torch.where(pos, SQRT2, 0.2 * SQRT2)
        y = av * torch.where(pos, 1 / SQRT2, 1 / (0.2 * SQRT2))
        if _addr(noise):
            y = y - T(noise_gain, (1,), torch.float32)[0] * TS(noise, (B, P), (noise_sb, 1), torch.float32).unsqueeze(-1)
        if _addr(bias):
            y = y - T(bias, (Cc,), torch.float32)
        if _addr(dd):
            T(dd, (B, Cc), torch.float32).add_((dz * y).sum(1))
        o = dz * mul
        if _addr(oscale):
            o = o * T(oscale, (B, Cc), torch.float32).unsqueeze(1)
        T(out, (B, P, Cc), torch.float16).copy_(o.clamp(-65504, 65504))
        return 0

    def b200ir_to_rgb_bwd(self, drgb, a, w, s, da, accumulate, ds, B, P, Cc, stream):
        self.launches += 1
        d = T(drgb, (B, 3, P), torch.float32)
        av = T(a, (B, P, Cc), torch.float16).float()
        t = torch.einsum('bop,oc->bpc', d, T(w, (3, Cc), torch.float32))
        if _addr(ds):
            T(ds, (B, Cc), torch.float32).add_((av * t).sum(1))
        o = t * (T(s, (B, Cc), torch.float32).unsqueeze(1) if _addr(s) else 1.0)
        if _addr(da):
            dt = T(da, (B, P, Cc), torch.float16)
            dt.copy_(((dt.float() + o) if accumulate else o).clamp(-65504, 65504))
        return 0

    def b200ir_style_act_bwd_params(self, da, a, noise, noise_sb, noise_gain, bias, oscale, mul, out, dd, db, dn, B, P, Cc, stream):
        dv, av = T(da, (B, P, Cc), torch.float16).float(), T(a, (B, P, Cc), torch.float16).float()
        dz = dv * torch.where(av > 0, SQRT2, 0.2 * SQRT2)
        T(db, (B, Cc), torch.float32).add_(dz.sum(1))
        if _addr(noise):
            T(dn, (B, Cc), torch.float32).add_((dz * TS(noise, (B, P), (noise_sb, 1), torch.float32).unsqueeze(-1)).sum(1))
        return self.b200ir_style_act_bwd(da, a, noise, noise_sb, noise_gain, bias, oscale, mul, out, dd, B, P, Cc, stream)

    def b200ir_to_rgb_bwd_params(self, drgb, a, w, s, da, accumulate, ds, R, B, P, Cc, stream):
        d = T(drgb, (B, 3, P), torch.float32)
        av = T(a, (B, P, Cc), torch.float16).float()
        T(R, (B, 3, Cc), torch.float32).add_(torch.einsum('bop,bpc->boc', d, av))
        return self.b200ir_to_rgb_bwd(drgb, a, w, s, da, accumulate, ds, B, P, Cc, stream)

    def b200ir_plane_sums(self, x, out, B, Cn, P, stream):
        self.launches += 1
        T(out, (Cn,), torch.float32).add_(T(x, (B, Cn, P), torch.float32).sum(dim=(0, 2)))
        return 0

    def b200ir_table_colsum(self, tab, in_f16, mul, m, scale, out, B, n, stream):
        self.launches += 1
        t = T(tab, (B, n), torch.float16 if in_f16 else torch.float32).float()
        if _addr(mul):
            t = (t.view(B, n // m, m) * T(mul, (B, 1, m), torch.float32)).view(B, n)
        T(out, (n,), torch.float32).copy_(scale * t.sum(0))
        return 0

    def b200ir_mod_linear_wgrad(self, ds, latent, wscale, dw, L, F_, lat_idx, B, cin, stream):
        self.launches += 1
        lat = T(latent, (B, L, F_), torch.float32)[:, lat_idx]
        T(dw, (cin, F_), torch.float32).copy_(wscale * T(ds, (B, cin), torch.float32).t() @ lat)
        return 0

    def b200ir_modconv_wgrad(self, G, transposed, W, s, dd, d, scale, dw, B, cin, cout, taps, stream):
        self.launches += 1
        if transposed:
            g = T(G, (cin, taps, cout), torch.float32).permute(2, 0, 1)
        else:
            g = T(G, (cout, taps, cin), torch.float32).permute(0, 2, 1)
        dv = T(d, (B, cout), torch.float32)
        t = T(dd, (B, cout), torch.float32) * dv * dv
        m = scale * scale * (t.t() @ T(s, (B, cin), torch.float32).pow(2))
        T(dw, (cout, cin, taps), torch.float32).copy_(scale * g - T(W, (cout, cin, taps), torch.float32) * m.unsqueeze(-1))
        return 0

    def b200ir_rgb_up_adjoint(self, d, out, planes, h, w, stream):
        self.launches += 1
        with torch.enable_grad():
            lo = torch.zeros(1, planes, h, w, requires_grad=True)
            upfirdn(lo, fir2d() * 4, up=2, pad=(2, 1)).backward(T(d, (1, planes, 2 * h, 2 * w), torch.float32))
        T(out, (planes, h, w), torch.float32).copy_(lo.grad[0])
        return 0

    def b200ir_demod_bwd(self, ds, s, dd, d, wsq, scale2, B, cin, cout, stream):
        self.launches += 1
        dv = T(d, (B, cout), torch.float32)
        t = T(dd, (B, cout), torch.float32) * dv * dv
        T(ds, (B, cin), torch.float32).sub_(scale2 * T(s, (B, cin), torch.float32) * (t @ T(wsq, (cout, cin), torch.float32)))
        return 0

    def b200ir_mod_linear_bwd(self, ds, w, wscale, dlat, L, Fd, lat_idx, B, cin, stream):
        self.launches += 1
        T(dlat, (B, L, Fd), torch.float32)[:, lat_idx].add_(wscale * T(ds, (B, cin), torch.float32) @ T(w, (cin, Fd), torch.float32))
        return 0

    def b200ir_first_conv_dgrad(self, dz, w, dx, accumulate, B, H, W, cout, stream):
        self.launches += 1
        g = torch.einsum('bhwc,ck->bkhw', T(dz, (B, H, W, cout), torch.float16).float(), T(w, (cout, 3), torch.float32))
        o = T(dx, (B, 3, H, W), torch.float32)
        o.copy_(o + g if accumulate else g)
        return 0

    def b200ir_head_to_nchw(self, head, rgb, B, P, cpad, stream):
        self.launches += 1
        T(rgb, (B, 3, P), torch.float32).copy_(T(head, (B, P, cpad), torch.float16)[..., :3].float().permute(0, 2, 1))
        return 0

    def b200ir_nchw_to_head(self, drgb, dhead, B, P, cpad, stream):
        self.launches += 1
        o = T(dhead, (B, P, cpad), torch.float16)
        o.zero_()
        o[..., :3].copy_(T(drgb, (B, 3, P), torch.float32).permute(0, 2, 1).clamp(-65504, 65504))
        return 0

    def b200ir_l1_loss(self, x, t, n, weight, grad_scale, loss, grad, stream):
        self.launches += 1
        d = T(x, (n,), torch.float32) - T(t, (n,), torch.float32)
        T(loss, (1,), torch.float32).add_(weight * d.abs().mean())
        if _addr(grad):
            T(grad, (n,), torch.float32).copy_(torch.sign(d) * (grad_scale * weight / n))
        return 0

    def b200ir_softplus_loss(self, pred, n, stride, sign, weight, grad_scale, loss, dpred, stream):
        self.launches += 1
        v = sign * TS(pred, (n,), (stride,), torch.float16).float()
        T(loss, (1,), torch.float32).add_(weight * F.softplus(v).mean())
        if _addr(dpred):
            TS(dpred, (n,), (stride,), torch.float16).copy_(grad_scale * weight / n * sign * torch.sigmoid(v))
        return 0

    def b200ir_nchw_to_nhwc_pad(self, x, out, B, Cc, H, W, Cpad, sub, mul, unshuffle, stream):
        self.launches += 1
        assert unshuffle == 1, 'cabi_sim: pixel_unshuffle variant not simulated'
        xt = T(x, (B, Cc, H, W), torch.float32)
        if _addr(sub):
            xt = xt - T(sub, (Cc,), torch.float32).view(1, Cc, 1, 1)
        o = T(out, (B, H, W, Cpad), torch.float16)
        o.zero_()
        o[..., :Cc].copy_((xt * mul).permute(0, 2, 3, 1).clamp(-65504, 65504))
        return 0

    def b200ir_gram_batched(self, x, dy, out, B, H, W, cin, cout, stream):
        self.launches += 1
        xt = T(x, (B, H * W, cin), torch.float16).float()
        dt = T(dy, (B, H * W, cout), torch.float16).float()
        T(out, (B, cout, cin), torch.float32).copy_(torch.einsum('bpo,bpi->boi', dt, xt))
        return 0

    def b200ir_maxpool2_relu(self, z, out, B, H, W, Cc, stream):
        self.launches += 1
        zt = nchw(T(z, (B, H, W, Cc), torch.float16).float())
        T(out, (B, H // 2, W // 2, Cc), torch.float16).copy_(nhwc(F.max_pool2d(F.relu(zt), 2, 2)))
        return 0

    def b200ir_maxpool2_relu_bwd(self, z, dpool, add, dz, B, H, W, Cc, stream):
        self.launches += 1
        o = torch.zeros(B, H, W, Cc)
        if _addr(dpool):
            with torch.enable_grad():
                zt = nchw(T(z, (B, H, W, Cc), torch.float16).float()).requires_grad_()
                F.max_pool2d(F.relu(zt), 2, 2).backward(nchw(T(dpool, (B, H // 2, W // 2, Cc), torch.float16).float()))
            o = nhwc(zt.grad)
        if _addr(add):
            o = o + T(add, (B, H, W, Cc), torch.float16).float()
        T(dz, (B, H, W, Cc), torch.float16).copy_(o.clamp(-65504, 65504))
        return 0

    def b200ir_l1_loss_f16(self, x, t, n, weight, grad_scale, loss, grad, stream):
        self.launches += 1
        d = T(x, (n,), torch.float16).float() - T(t, (n,), torch.float16).float()
        T(loss, (1,), torch.float32).add_(weight * d.abs().mean())
        if _addr(grad):
            T(grad, (n,), torch.float16).copy_(torch.sign(d) * (grad_scale * weight / n))
        return 0

    def b200ir_sum_squares(self, x, n, scale, out, stream):
        self.launches += 1
        T(out, (1,), torch.float32).add_(scale * T(x, (n,), torch.float32).pow(2).sum())
        return 0

    @staticmethod
    def _mbstd_stat(x, group):
        """(a-weighted) statistic of minibatch_stddev as a differentiable torch function of x [B,P,C]: returns s [M]."""
        B, P, Cc = x.shape
        g = x.view(group, B // group, P, Cc)
        return torch.sqrt(g.var(0, unbiased=False) + 1e-8).mean(dim=(1, 2))

    def b200ir_minibatch_stddev_jvp(self, x, t, ts, tcat, B, P, Cc, c_pad, group, stream):
        self.launches += 1
        xt, tt = T(x, (B, P, Cc), torch.float16).float(), T(t, (B, P, Cc), torch.float16).float()
        _, jv = torch.autograd.functional.jvp(lambda v: self._mbstd_stat(v, group), xt, tt)
        T(ts, (B // group,), torch.float32).copy_(jv)
        o = T(tcat, (B, P, c_pad), torch.float16)
        o.zero_()
        o[..., :Cc].copy_(tt)
        o[..., Cc].copy_(jv.repeat(group).view(B, 1).expand(B, P))
        return 0

    def b200ir_minibatch_stddev_hvp(self, x, t, a, q, B, P, Cc, group, stream):
        self.launches += 1
        xt, tt = T(x, (B, P, Cc), torch.float16).float(), T(t, (B, P, Cc), torch.float16).float()
        av = T(a, (B // group,), torch.float32)
        _, hv = torch.autograd.functional.hvp(lambda v: (self._mbstd_stat(v, group) * av).sum(), xt, tt)
        T(q, (B, P, Cc), torch.float16).copy_(hv.clamp(-65504, 65504))
        return 0

    def b200ir_pack_weights(self, w, out, cout, cin, kh, kw, scale, mode, cin_pad, stream):
        self.launches += 1
        wt = T(w, (cout, cin, kh, kw), torch.float32) * scale
        if mode == 0:
            cp = cin_pad or cin
            o = T(out, (cout, kh * kw, cp), torch.float16)
            o.zero_()
            o[..., :cin].copy_(wt.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin))
        else:
            T(out, (cin, kh * kw * cout), torch.float16).copy_(wt.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, kh * kw * cout))
        return 0


class installed:
    """Context manager: routes image_restoration_b200 through the simulator (and lifts the CUDA-only guards) for host-logic
    tests on CPU tensors."""

    def __enter__(self):
        import contextlib
        from image_restoration_b200 import _lib, ops
        self.saved = (_lib._lib, _lib.require_cuda, _lib.device_ctx, ops._stream, ops._PLANS.plans)
        self.sim = SimLib()
        _lib._lib = self.sim
        _lib.require_cuda = lambda t, who: None
        _lib.device_ctx = lambda dev: contextlib.nullcontext()
        ops._stream = lambda: None
        ops._PLANS.plans = {}
        return self.sim

    def __exit__(self, *exc):
        from image_restoration_b200 import _lib, ops
        _lib._lib, _lib.require_cuda, _lib.device_ctx, ops._stream, ops._PLANS.plans = self.saved
        return False
