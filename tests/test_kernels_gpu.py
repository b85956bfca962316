"""GPU parity tests of the individual kernels (through the C ABI) against plain torch fp32 restatements of the
reference operators, computed from the same fp16-rounded inputs.  Tolerances: outputs are stored in fp16
(rel 2^-11) after fp32 accumulation, so |err| <= 4e-3 * scale is expected; asserted at 1e-2 * scale."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def setup_module(module):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _ops():
    from image_restoration_b200 import ops
    return ops


def nhwc16(t):  # NCHW fp32 -> NHWC fp16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().half()


def nchw32(t):  # NHWC fp16 -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def pack3x3(w):  # (cout,cin,kh,kw) -> [cout][(kh*kw)*cin]
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().half()


def check_close(got, ref, tol=1e-2, what=''):
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    print(f'{what}: max|err|={err:.3e} scale={scale:.3e} rel={err / scale:.3e}')
    assert math.isfinite(err) and err <= tol * scale, f'{what}: err {err} vs scale {scale}'


def fir_k(dev):
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = k[None] * k[:, None]
    return k / k.sum()


def upfirdn_ref(x, k, up=1, down=1, pad=(0, 0)):
    b, c, h, w = x.shape
    if up > 1:
        z = x.new_zeros(b, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.view(b, c, h * up, w * up)
    x = F.pad(x, [pad[0], pad[1], pad[0], pad[1]])
    y = F.conv2d(x.reshape(b * c, 1, *x.shape[2:]), torch.flip(k, [0, 1])[None, None])
    return y.view(b, c, *y.shape[2:])[:, :, ::down, ::down]


@pytest.mark.parametrize('B,H,W,cin,cout,k', [
    (2, 32, 96, 256, 256, 3),
    (1, 128, 384, 32, 32, 3),
    (3, 16, 48, 512, 512, 3),
    (2, 64, 192, 64, 64, 3),
    (5, 4, 12, 256, 256, 3),
    (3, 8, 24, 512, 512, 3),
    (2, 32, 96, 256, 64, 1),
    (2, 64, 192, 128, 128, 3),
    (1, 32, 32, 16, 16, 3),
])
def test_conv_same_bias_act(B, H, W, cin, cout, k):
    ops = _ops()
    torch.manual_seed(0)
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, k, k, device=DEV) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, device=DEV) * 0.1
    xh, wh = nhwc16(x), pack3x3(w)
    out = torch.empty(B, H, W, cout, device=DEV, dtype=torch.float16)
    op = ops.conv_same(xh, wh, out, k, bias=bias, act=True)
    op()
    torch.cuda.synchronize()
    wr = wh.float().view(cout, k, k, cin).permute(0, 3, 1, 2)
    ref = F.leaky_relu(F.conv2d(nchw32(xh), wr, bias, padding=k // 2), 0.2) * math.sqrt(2)
    check_close(nchw32(out), ref, what=f'conv{k}x{k} {cin}->{cout} @{H}x{W} B{B}')


def test_conv_residual_and_channel_offset():
    ops = _ops()
    torch.manual_seed(1)
    B, H, W, cin, cout = 2, 16, 48, 256, 256
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    res = torch.randn(B, cout, H, W, device=DEV)
    xh, wh, rh = nhwc16(x), pack3x3(w), nhwc16(res)
    out = torch.zeros(B, H, W, 2 * cout, device=DEV, dtype=torch.float16)
    ops.conv_same(xh, wh, out, 3, out_c_off=cout, res=rh, res_mode=1,
                  res_strides=(cout, W * cout, H * W * cout), res_wh=(W, H), res_scale=1 / math.sqrt(2))()
    torch.cuda.synchronize()
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = (F.conv2d(nchw32(xh), wr, padding=1) + nchw32(rh)) / math.sqrt(2)
    got = nchw32(out)
    check_close(got[:, cout:], ref, what='conv+res into channel slice')
    assert got[:, :cout].abs().max().item() == 0


def test_conv_bilinear_residual():
    ops = _ops()
    torch.manual_seed(2)
    B, H, W, cin, cout = 2, 16, 48, 256, 64
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    lo = torch.randn(B, cout, H // 2, W // 2, device=DEV)
    xh, wh, lh = nhwc16(x), pack3x3(w), nhwc16(lo)
    out = torch.empty(B, H, W, cout, device=DEV, dtype=torch.float16)
    ops.conv_same(xh, wh, out, 3, res=lh, res_mode=2, res_strides=(cout, (W // 2) * cout, (H // 2) * (W // 2) * cout),
                  res_wh=(W // 2, H // 2), res_scale=1 / math.sqrt(2))()
    torch.cuda.synchronize()
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    up = F.interpolate(nchw32(lh), scale_factor=2, mode='bilinear', align_corners=False)
    ref = (F.conv2d(nchw32(xh), wr, padding=1) + up) / math.sqrt(2)
    check_close(nchw32(out), ref, what='conv+bilinear residual')


def test_modulated_epilogue_demod_noise():
    ops = _ops()
    torch.manual_seed(3)
    B, H, W, cin, cout = 3, 8, 24, 512, 512
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    demod = torch.rand(B, cout, device=DEV) + 0.5
    noise = torch.randn(1, 1, H, W, device=DEV)
    gain = torch.tensor([0.3], device=DEV)
    bias = torch.randn(cout, device=DEV) * 0.1
    xh, wh = nhwc16(x), pack3x3(w)
    out = torch.empty(B, H, W, cout, device=DEV, dtype=torch.float16)
    ops.conv_same(xh, wh, out, 3, bias=bias, demod=demod, noise=noise, noise_gain=gain, noise_strides=(0, W), act=True)()
    torch.cuda.synchronize()
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    y = F.conv2d(nchw32(xh), wr, padding=1) * demod[:, :, None, None] + gain * noise + bias[None, :, None, None]
    ref = F.leaky_relu(y, 0.2) * math.sqrt(2)
    check_close(nchw32(out), ref, what='modconv epilogue')


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 32, 96, 64, 256), (1, 128, 384, 32, 64), (3, 8, 24, 256, 256)])
def test_fir_then_stride2_conv(B, H, W, cin, cout):
    ops = _ops()
    torch.manual_seed(4)
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    bias = torch.randn(cout, device=DEV) * 0.1
    xh, wh = nhwc16(x), pack3x3(w)
    p = torch.zeros(B, H + 2, W + 2, cin, device=DEV, dtype=torch.float16)
    ops.fir_pad22(xh, p)
    torch.cuda.synchronize()
    pref = upfirdn_ref(nchw32(xh), fir_k(DEV), pad=(2, 2))
    check_close(nchw32(p)[:, :, :H + 1, :W + 1], pref, what='fir_pad22')
    out = torch.empty(B, H // 2, W // 2, cout, device=DEV, dtype=torch.float16)
    ops.conv3x3_s2(p, H, W, wh, out, bias=bias, act=True)()
    torch.cuda.synchronize()
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = F.leaky_relu(F.conv2d(nchw32(p)[:, :, :H + 1, :W + 1], wr, bias, stride=2), 0.2) * math.sqrt(2)
    check_close(nchw32(out), ref, what=f'conv3x3 s2 {cin}->{cout} @{H}x{W}')


def test_fir_down2_and_bilinear_and_add():
    ops = _ops()
    torch.manual_seed(5)
    B, H, W, C = 2, 32, 96, 64
    xh = nhwc16(torch.randn(B, C, H, W, device=DEV))
    out = torch.empty(B, H // 2, W // 2, C, device=DEV, dtype=torch.float16)
    ops.fir_down2(xh, out)
    ref = upfirdn_ref(nchw32(xh), fir_k(DEV), pad=(1, 1))[:, :, ::2, ::2]
    check_close(nchw32(out), ref, what='fir_down2')
    up = torch.empty(B, 2 * H, 2 * W, C, device=DEV, dtype=torch.float16)
    ops.bilinear_up2(xh, up)
    check_close(nchw32(up), F.interpolate(nchw32(xh), scale_factor=2, mode='bilinear', align_corners=False),
                what='bilinear_up2')
    s = torch.empty_like(xh)
    ops.add(xh, xh, s)
    check_close(s.float(), 2 * xh.float(), what='add')


@pytest.mark.parametrize('B,h,w,cin,cout,sft', [(2, 4, 12, 512, 512, True), (2, 32, 96, 512, 128, True),
                                                 (1, 64, 192, 128, 64, False)])
def test_transposed_conv_phases_and_tail(B, h, w, cin, cout, sft):
    ops = _ops()
    torch.manual_seed(6)
    x = torch.randn(B, cin, h, w, device=DEV)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    demod = torch.rand(B, cout, device=DEV) + 0.5
    xh = nhwc16(x)
    wh = wt.half()
    raw = torch.zeros(B, 2 * h + 2, 2 * w + 2, cout, device=DEV, dtype=torch.float16)
    for (py, px) in ops.CONVT_PHASES:
        taps = ops.convt_phase_taps(py, px)
        wp = torch.cat([wh[:, :, kh, kw] for kh, kw in taps], dim=1).contiguous()
        ops.convt_s2_phase(xh, wp, py, px, raw, demod=demod)()
    torch.cuda.synchronize()
    ref_raw = F.conv_transpose2d(nchw32(xh), wh.float().transpose(0, 1), stride=2) * demod[:, :, None, None]
    check_close(nchw32(raw)[:, :, :2 * h + 1, :2 * w + 1], ref_raw, what='convT phases')
    # tail
    h2, w2 = 2 * h, 2 * w
    noise = torch.randn(B, 1, h2, w2, device=DEV)
    gain = torch.tensor([0.2], device=DEV)
    bias = torch.randn(cout, device=DEV) * 0.1
    c_sft = cout // 2 if sft else cout
    scale = nhwc16(torch.randn(B, c_sft, h2, w2, device=DEV))
    shift = nhwc16(torch.randn(B, c_sft, h2, w2, device=DEV))
    s_next = torch.rand(B, cout, device=DEV) + 0.5
    out = torch.empty(B, h2, w2, cout, device=DEV, dtype=torch.float16)
    ops.upfir_act(raw, out, noise, h2 * w2, gain, bias, scale, shift, c_sft, s_next)
    torch.cuda.synchronize()
    y = upfirdn_ref(nchw32(raw)[:, :, :2 * h + 1, :2 * w + 1], fir_k(DEV) * 4, pad=(1, 1))
    y = F.leaky_relu(y + gain * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    keep = cout - c_sft
    y = torch.cat([y[:, :keep], y[:, keep:] * nchw32(scale) + nchw32(shift)], 1) * s_next[:, :, None, None]
    check_close(nchw32(out), y, what='upfir_act')


def test_linear_as_conv():
    ops = _ops()
    torch.manual_seed(7)
    B, K, N = 5, 12288, 3072
    x = (torch.randn(B, K, device=DEV)).half()
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).half()
    bias = torch.randn(N, device=DEV)
    out = torch.empty(B, N, device=DEV, dtype=torch.float32)
    ops.linear_as_conv(x, w, out, bias=bias, block_n=64)()
    torch.cuda.synchronize()
    check_close(out, x.float() @ w.float().t() + bias, tol=2e-3, what='linear 12288->3072')


def test_style_path_and_to_rgb():
    ops = _ops()
    torch.manual_seed(8)
    B, L, Fd, cin, cout = 3, 12, 256, 512, 128
    latent = torch.randn(B, L, Fd, device=DEV)
    wm = torch.randn(cin, Fd, device=DEV)
    bm = torch.ones(cin, device=DEV)
    s = torch.empty(B, cin, device=DEV)
    ops.mod_linear(latent, 5, wm, bm, 1 / math.sqrt(Fd), s)
    sref = latent[:, 5] @ wm.t() / math.sqrt(Fd) + bm
    check_close(s, sref, tol=1e-5, what='mod_linear')
    wconv = torch.randn(cout, cin, 3, 3, device=DEV)
    wsq = wconv.pow(2).sum([2, 3]).contiguous()
    d = torch.empty(B, cout, device=DEV)
    scale2 = 1.0 / (cin * 9)
    ops.demod(s, wsq, scale2, d)
    wmod = (1 / math.sqrt(cin * 9)) * wconv[None] * sref[:, None, :, None, None]
    dref = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
    check_close(d, dref, tol=1e-4, what='demod')
    cst = nhwc16(torch.randn(1, cin, 4, 12, device=DEV))
    xs = torch.empty(B, 4, 12, cin, device=DEV, dtype=torch.float16)
    ops.modulate_const(cst, s, xs)
    check_close(nchw32(xs), nchw32(cst) * sref[:, :, None, None], what='modulate_const')
    # to_rgb with skip upsample and xs by-product
    h, w, C = 16, 48, 128
    x = nhwc16(torch.randn(B, C, h, w, device=DEV))
    wrgb = torch.randn(3, C, device=DEV) / math.sqrt(C)
    srgb = torch.rand(B, C, device=DEV) + 0.5
    snext = torch.rand(B, C, device=DEV) + 0.5
    brgb = torch.randn(3, device=DEV)
    skip = torch.randn(B, 3, h // 2, w // 2, device=DEV)
    rgb = torch.empty(B, 3, h, w, device=DEV)
    xs2 = torch.empty_like(x)
    ops.to_rgb(x, wrgb, srgb, brgb, skip, rgb, snext, xs2)
    torch.cuda.synchronize()
    ref = torch.einsum('bchw,oc,bc->bohw', nchw32(x), wrgb, srgb) + brgb[None, :, None, None]
    ref = ref + upfirdn_ref(skip, fir_k(DEV) * 4, up=2, pad=(2, 1))
    check_close(rgb, ref, tol=1e-4, what='to_rgb')
    check_close(nchw32(xs2), nchw32(x) * snext[:, :, None, None], what='to_rgb xs by-product')
    for Cc in (64, 512):
        x = nhwc16(torch.randn(B, Cc, 8, 24, device=DEV))
        wr = torch.randn(3, Cc, device=DEV) / math.sqrt(Cc)
        rgb = torch.empty(B, 3, 8, 24, device=DEV)
        ops.to_rgb(x, wr, None, brgb, None, rgb)
        check_close(rgb, torch.einsum('bchw,oc->bohw', nchw32(x), wr) + brgb[None, :, None, None], tol=1e-4,
                    what=f'plain toRGB C={Cc}')


def test_first_conv():
    ops = _ops()
    torch.manual_seed(9)
    B, H, W, cout = 2, 128, 384, 32
    x = torch.rand(B, 3, H, W, device=DEV) * 2 - 1
    w = torch.randn(cout, 3, device=DEV) / math.sqrt(3)
    bias = torch.randn(cout, device=DEV) * 0.1
    out = torch.empty(B, H, W, cout, device=DEV, dtype=torch.float16)
    ops.first_conv(x, w, bias, out)
    ref = F.leaky_relu(F.conv2d(x, w[:, :, None, None], bias), 0.2) * math.sqrt(2)
    check_close(nchw32(out), ref, what='first_conv')


def test_style_path_multi_launch():
    """All modulation linears / demod tables of a forward in one launch each == the per-layer kernels."""
    ops = _ops()
    torch.manual_seed(10)
    B, L, Fd = 5, 12, 256
    latent = torch.randn(B, L, Fd, device=DEV)
    dims = [(512, 512), (512, 3), (128, 64), (64, 64)]
    mods, dems, refs = [], [], []
    for j, (cin, cout) in enumerate(dims):
        wm, bm = torch.randn(cin, Fd, device=DEV), torch.randn(cin, device=DEV)
        s = torch.empty(B, cin, device=DEV)
        wsq = torch.rand(cout, cin, device=DEV) * 9
        d = torch.empty(B, cout, device=DEV)
        mods.append((wm, bm, j + 2, s))
        dems.append((s, wsq, 1.0 / (cin * 9), d))
        sref = latent[:, j + 2] @ wm.t() / math.sqrt(Fd) + bm
        refs.append((sref, torch.rsqrt(sref.pow(2) @ wsq.t() / (cin * 9) + 1e-8)))
    ops.ModLinearMulti(latent, mods, 1 / math.sqrt(Fd))()
    ops.DemodMulti(dems)()
    torch.cuda.synchronize()
    for (wm, bm, li, s), (_, _, _, d), (sref, dref) in zip(mods, dems, refs):
        check_close(s, sref, tol=1e-5, what='mod_linear_multi')
        check_close(d, dref, tol=1e-4, what='demod_multi')


@pytest.mark.parametrize('B,h2,w2,C,c_sft,with_noise', [
    (3, 128, 384, 64, 32, True),     # CC=64 chunk straddles the SFT boundary (32 SFT channels per pixel)
    (2, 64, 192, 128, 64, True),
    (2, 32, 96, 512, 256, False),    # OW=96: one output per thread variant
    (5, 8, 24, 512, 256, True),
    (2, 16, 48, 32, 16, True),       # 32-channel chunks
    (2, 40, 72, 64, 64, True),       # sft_half=False: every channel modulated; ragged strips / row chunks
    (1, 16, 16, 96, 0, True),        # no SFT, C % 64 != 0
])
def test_upfir_act_streaming(B, h2, w2, C, c_sft, with_noise):
    """TMA-fed streaming FIR tail vs upfirdn2d restatement; the unused last row/col of `raw` is poisoned to prove the
    kernel reads only the (h2+1)x(w2+1) valid samples (zero padding comes from the TMA out-of-bounds fill)."""
    ops = _ops()
    torch.manual_seed(11)
    raw = nhwc16(torch.randn(B, C, h2 + 2, w2 + 2, device=DEV))
    raw[:, h2 + 1, :, :] = 1000.0
    raw[:, :, w2 + 1, :] = -1000.0
    noise = torch.randn(B, 1, h2, w2, device=DEV) if with_noise else None
    gain = torch.tensor([0.3], device=DEV)
    bias = torch.randn(C, device=DEV) * 0.1
    scale = nhwc16(torch.randn(B, c_sft, h2, w2, device=DEV)) if c_sft else None
    shift = nhwc16(torch.randn(B, c_sft, h2, w2, device=DEV)) if c_sft else None
    s_next = torch.rand(B, C, device=DEV) + 0.5
    out = torch.empty(B, h2, w2, C, device=DEV, dtype=torch.float16)
    ops.upfir_act(raw, out, noise, h2 * w2, gain, bias, scale, shift, c_sft, s_next)
    torch.cuda.synchronize()
    y = upfirdn_ref(nchw32(raw)[:, :, :h2 + 1, :w2 + 1], fir_k(DEV) * 4, pad=(1, 1))
    y = y + bias[None, :, None, None]
    if with_noise:
        y = y + gain * noise
    y = F.leaky_relu(y, 0.2) * math.sqrt(2)
    keep = C - c_sft
    if c_sft:
        y = torch.cat([y[:, :keep], y[:, keep:] * nchw32(scale) + nchw32(shift)], 1)
    y = y * s_next[:, :, None, None]
    check_close(nchw32(out), y, what=f'upfir_act streaming C={C} {h2}x{w2}')


@pytest.mark.parametrize('B,H,W,C', [(2, 128, 384, 32), (3, 64, 192, 64), (2, 32, 96, 256), (4, 8, 24, 256),
                                     (1, 20, 36, 64), (2, 16, 16, 16)])
def test_fir_pad22_streaming(B, H, W, C):
    ops = _ops()
    torch.manual_seed(12)
    xh = nhwc16(torch.randn(B, C, H, W, device=DEV))
    p = torch.full((B, H + 2, W + 2, C), 7.0, device=DEV, dtype=torch.float16)
    ops.fir_pad22(xh, p)
    torch.cuda.synchronize()
    pref = upfirdn_ref(nchw32(xh), fir_k(DEV), pad=(2, 2))
    got = nchw32(p)
    check_close(got[:, :, :H + 1, :W + 1], pref, what=f'fir_pad22 streaming C={C} {H}x{W}')
    assert (got[:, :, H + 1] == 7.0).all() and (got[:, :, :, W + 1] == 7.0).all()   # nothing written past (H+1)x(W+1)


@pytest.mark.parametrize('B,H,W,C,no_store,with_skip', [(2, 16, 48, 512, False, True), (3, 64, 192, 64, True, True),
                                                        (64, 4, 12, 512, False, False), (2, 128, 384, 64, True, True)])
def test_conv_fused_to_rgb_and_next_modulation(B, H, W, C, no_store, with_skip):
    """StyleConv (plain) + ToRGB fused: conv epilogue writes out * s_next and the three ToRGB dot products as partial
    planes; rgb_combine adds bias, partials and the up-sampled skip (stylegan2_ocr_arch.py:323-333,357-374)."""
    ops = _ops()
    torch.manual_seed(13)
    x = torch.randn(B, C, H, W, device=DEV)
    w = torch.randn(C, C, 3, 3, device=DEV) / math.sqrt(C * 9)
    xh, wh = nhwc16(x), pack3x3(w)
    bias = torch.randn(C, device=DEV) * 0.1
    demod = torch.rand(B, C, device=DEV) + 0.5
    noise = torch.randn(B, 1, H, W, device=DEV)
    gain = torch.tensor([0.1], device=DEV)
    s_next = torch.rand(B, C, device=DEV) + 0.5
    wrgb = torch.randn(3, C, device=DEV) / math.sqrt(C)
    s_rgb = torch.rand(B, C, device=DEV) + 0.5
    brgb = torch.randn(3, device=DEV)
    wm = torch.empty(B, 3, C, device=DEV)
    ops.rgb_wmod(wrgb, s_rgb, wm)
    out = None if no_store else torch.empty(B, H, W, C, device=DEV, dtype=torch.float16)
    op = ops.conv_same(xh, wh, out, 3, bias=bias, demod=demod, noise=noise, noise_gain=gain,
                       noise_strides=(H * W, W), act=True, out_scale=s_next)
    part = op.attach_rgb(wm, (H, W), no_store=no_store)
    op()
    skip = torch.randn(B, 3, H // 2, W // 2, device=DEV) if with_skip else None
    rgb = torch.empty(B, 3, H, W, device=DEV)
    ops.rgb_combine(part, brgb, skip, rgb)
    torch.cuda.synchronize()
    wr = wh.float().view(C, 3, 3, C).permute(0, 3, 1, 2)
    y = F.conv2d(nchw32(xh), wr, padding=1) * demod[:, :, None, None] + gain * noise + bias[None, :, None, None]
    y = F.leaky_relu(y, 0.2) * math.sqrt(2)
    ref_rgb = torch.einsum('bchw,oc,bc->bohw', y, wrgb, s_rgb) + brgb[None, :, None, None]
    if with_skip:
        ref_rgb = ref_rgb + upfirdn_ref(skip, fir_k(DEV) * 4, up=2, pad=(2, 1))
    check_close(rgb, ref_rgb, tol=2e-3, what=f'fused toRGB C={C} {H}x{W}')
    if not no_store:
        check_close(nchw32(out), y * s_next[:, :, None, None], what='conv out * s_next')


@pytest.mark.parametrize('B,H,W,cin,cout,kind', [
    (1, 128, 384, 32, 32, 'plain'), (2, 64, 192, 64, 64, 'plain'), (3, 40, 200, 32, 64, 'plain'),
    (1, 128, 384, 64, 32, 'res2'), (2, 16, 130, 16, 16, 'plain'), (2, 9, 128, 64, 64, 'modrgb'),
    (2, 33, 256, 64, 64, 'modrgb'), (1, 1, 128, 32, 32, 'plain'), (2, 2, 140, 64, 32, 'plain'),
    (5, 24, 384, 128, 64, 'plain'), (2, 40, 256, 32, 32, 'res1'), (3, 17, 130, 32, 16, 'res1'),
])
def test_row_sliding_conv_forced(B, H, W, cin, cout, kind):
    """The row-sliding conv variant (one MMA of N = 3*cout per input row and kw tap, accumulator ring in TMEM), forced
    at small sizes so that ring wrap-around, short row chunks and ragged widths are covered."""
    ops = _ops()
    torch.manual_seed(14)
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    bias = torch.randn(cout, device=DEV) * 0.1
    xh, wh = nhwc16(x), pack3x3(w)
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    out = torch.empty(B, H, W, cout, device=DEV, dtype=torch.float16)
    kw = dict(bias=bias, act=True, tile=(128, 1, 1), row_mode=2, block_n=cout)
    y = F.conv2d(nchw32(xh), wr, padding=1)
    if kind == 'res2':
        assert H % 2 == 0 and W % 2 == 0
        lo = nhwc16(torch.randn(B, cout, H // 2, W // 2, device=DEV))
        kw.update(res=lo, res_mode=2, res_strides=(cout, (W // 2) * cout, (H // 2) * (W // 2) * cout),
                  res_wh=(W // 2, H // 2), res_scale=1 / math.sqrt(2))
        ref = (F.leaky_relu(y + bias[None, :, None, None], 0.2) * math.sqrt(2) +
               F.interpolate(nchw32(lo), scale_factor=2, mode='bilinear', align_corners=False)) / math.sqrt(2)
    elif kind == 'res1':      # same-resolution residual (the two-CTAs-per-SM instantiation of the residual profile)
        rr = nhwc16(torch.randn(B, cout, H, W, device=DEV))
        kw.update(res=rr, res_mode=1, res_strides=(cout, W * cout, H * W * cout), res_wh=(W, H), res_scale=1 / math.sqrt(2))
        ref = (F.leaky_relu(y + bias[None, :, None, None], 0.2) * math.sqrt(2) + nchw32(rr)) / math.sqrt(2)
    elif kind == 'modrgb':
        demod = torch.rand(B, cout, device=DEV) + 0.5
        noise = torch.randn(B, 1, H, W, device=DEV)
        gain = torch.tensor([0.1], device=DEV)
        s_next = torch.rand(B, cout, device=DEV) + 0.5
        wm = torch.randn(B, 3, cout, device=DEV) / math.sqrt(cout)
        kw.update(demod=demod, noise=noise, noise_gain=gain, noise_strides=(H * W, W), out_scale=s_next)
        yy = F.leaky_relu(y * demod[:, :, None, None] + gain * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
        ref = yy * s_next[:, :, None, None]
    else:
        ref = F.leaky_relu(y + bias[None, :, None, None], 0.2) * math.sqrt(2)
    op = ops.conv_same(xh, wh, out, 3, **kw)
    part = op.attach_rgb(wm, (H, W)) if kind == 'modrgb' else None
    op()
    op()     # a second launch must give the same result (no state left in TMEM / barriers)
    torch.cuda.synchronize()
    check_close(nchw32(out), ref, what=f'row conv {cin}->{cout} @{H}x{W} B{B} {kind}')
    if kind == 'modrgb':
        check_close(part.sum(0), torch.einsum('bchw,boc->bohw', yy, wm), tol=2e-3, what='row conv fused rgb')


@pytest.mark.parametrize('B,h,w,cin,cout', [(2, 4, 12, 512, 512), (3, 32, 96, 512, 128), (2, 64, 192, 128, 64),
                                            (1, 5, 7, 64, 64), (9, 8, 24, 512, 512)])
def test_transposed_conv_merged_single_gemm(B, h, w, cin, cout):
    """conv_transpose2d(stride 2) as ONE implicit GEMM (phases = column blocks, per-N-tile tap masks, phase-scatter
    store) == F.conv_transpose2d * demod on the (2h+1) x (2w+1) valid region; the padding row / column may hold anything."""
    ops = _ops()
    torch.manual_seed(15)
    x = torch.randn(B, cin, h, w, device=DEV)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    demod = torch.rand(B, cout, device=DEV) + 0.5
    xh = nhwc16(x)
    w_big = ops.convt_merged_weight(wt, 1.0)
    raw = torch.full((B, 2 * h + 2, 2 * w + 2, cout), 3.0, device=DEV, dtype=torch.float16)
    op = ops.convt_s2_merged(xh, w_big, raw, demod)
    op()
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(nchw32(xh), wt.half().float().transpose(0, 1), stride=2) * demod[:, :, None, None]
    check_close(nchw32(raw)[:, :, :2 * h + 1, :2 * w + 1], ref, what=f'merged convT {cin}->{cout} {h}x{w}')


@pytest.mark.parametrize('B,H,W,C', [(2, 128, 384, 32), (3, 64, 192, 64), (2, 32, 96, 256), (5, 8, 24, 256),
                                     (64, 4, 12, 256), (1, 34, 70, 64), (2, 6, 10, 96), (2, 16, 16, 16)])
def test_resamplers_streaming(B, H, W, C):
    """fir_down2 (FIR pad (1,1) + stride-2 sampling) and bilinear_up2 through the TMA-fed streaming kernels (C % 32 == 0)
    and the direct kernels (C = 16), ragged strips and row chunks included."""
    ops = _ops()
    torch.manual_seed(16)
    xh = nhwc16(torch.randn(B, C, H, W, device=DEV))
    down = torch.full((B, H // 2, W // 2, C), 9.0, device=DEV, dtype=torch.float16)
    ops.fir_down2(xh, down)
    ref = upfirdn_ref(nchw32(xh), fir_k(DEV), pad=(1, 1))[:, :, ::2, ::2]
    check_close(nchw32(down), ref, what=f'fir_down2 C={C} {H}x{W}')
    up = torch.full((B, 2 * H, 2 * W, C), 9.0, device=DEV, dtype=torch.float16)
    ops.bilinear_up2(xh, up)
    torch.cuda.synchronize()
    check_close(nchw32(up), F.interpolate(nchw32(xh), scale_factor=2, mode='bilinear', align_corners=False),
                what=f'bilinear_up2 C={C} {H}x{W}')


@pytest.mark.parametrize('B,h,w,cin,cout', [(2, 8, 24, 256, 256), (2, 32, 96, 256, 64), (1, 64, 192, 64, 32),
                                            (3, 5, 7, 64, 64), (2, 4, 12, 256, 256)])
def test_conv_up_layer_folded(B, h, w, cin, cout):
    """ResUpBlock tail with the ConvUpLayer folded (ops.UpFoldConv): one conv over the replicate-padded low-resolution
    input with the four output phases as column blocks + border-ring corrections ==
    (lrelu(conv3x3(F.interpolate(t, 2, bilinear)) + b) * sqrt2 + F.interpolate(skip, 2, bilinear)) / sqrt2."""
    ops = _ops()
    torch.manual_seed(17)
    t = torch.randn(B, cin, h, w, device=DEV)
    wt = torch.randn(cout, cin, 3, 3, device=DEV)
    scale = 1 / math.sqrt(cin * 9)
    bias = torch.randn(cout, device=DEV) * 0.2
    sl = nhwc16(torch.randn(B, cout, h, w, device=DEV))
    th = nhwc16(t)
    tp = torch.zeros(B, h + 2, w + 2, cin, device=DEV, dtype=torch.float16)
    tp[:, 1:-1, 1:-1] = th
    out = torch.empty(B, 2 * h, 2 * w, cout, device=DEV, dtype=torch.float16)
    uf = ops.UpFoldConv(tp, ops.upfold_weights(wt, scale), bias, out, sl, 1 / math.sqrt(2))
    uf.pad()
    for op in uf.border_ops:
        op()
    uf.corners()
    uf.main()
    torch.cuda.synchronize()
    up = lambda z: F.interpolate(z, scale_factor=2, mode='bilinear', align_corners=False)   # noqa: E731
    y = F.conv2d(up(nchw32(th)), (wt * scale).half().float(), bias, padding=1)
    ref = (F.leaky_relu(y, 0.2) * math.sqrt(2) + up(nchw32(sl))) / math.sqrt(2)
    got = nchw32(out)
    err_in = (got - ref)[:, :, 1:-1, 1:-1].abs().max().item() if h > 1 else 0.0
    err = (got - ref).abs().max().item()
    print(f'folded ConvUpLayer {cin}->{cout} {h}x{w}: interior {err_in:.3e} full {err:.3e} scale {ref.abs().max().item():.2f}')
    assert err <= 1.5e-2 * ref.abs().max().item()


@pytest.mark.parametrize('cin,cout,H,W', [(256, 256, 16, 48), (32, 32, 128, 384)])
def test_fp16_stores_saturate_instead_of_overflowing(cin, cout, H, W):
    """Activations are stored as fp16; the reference computes in fp32.  Values beyond +-65504 must clamp to the largest
    finite fp16 (F2FP.SATFINITE) rather than become inf, which would turn into NaN in the next layer."""
    ops = _ops()
    torch.manual_seed(3)
    x = torch.randn(1, cin, H, W, device=DEV) * 1000
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 2
    xh, wh = nhwc16(x), pack3x3(w)
    out = torch.empty(1, H, W, cout, device=DEV, dtype=torch.float16)
    ops.conv_same(xh, wh, out, 3, bias=torch.zeros(cout, device=DEV), act=True)()
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(nchw32(xh), wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), padding=1), 0.2) * math.sqrt(2)
    assert ref.abs().max().item() > 65504                     # the case is real: fp32 result exceeds the fp16 range
    got = nchw32(out)
    assert torch.isfinite(got).all()
    assert torch.equal(got, ref.clamp(-65504, 65504).half().float()) or \
        (got - ref.clamp(-65504, 65504)).abs().max().item() <= 1e-2 * 65504
    assert (got.abs() == 65504).any()


@pytest.mark.parametrize('B,H,W,cin,cout,kind', [(8, 32, 96, 256, 256, 'plain'), (7, 36, 96, 256, 512, 'plain'),
                                                 (4, 64, 192, 128, 128, 'mod'), (6, 32, 96, 128, 64, 'convt'),
                                                 (8, 32, 96, 256, 256, 'res')])
def test_conv_cta_pair_matches_single_cta(B, H, W, cin, cout, kind):
    """Layers with streamed weights, N-tiles of 128 / 256 columns and at least one M-tile per SM run on CTA pairs
    (tcgen05.mma.cta_group::2, M = 256: each CTA stages its own A rows and half of the weight tile).  Same K order per
    accumulator, so the result must equal the single-CTA kernel's bit for bit (max_ctas > 0 keeps a launch on single CTAs);
    (7, 36, 96) has an odd number of M-tiles: the last pair works with an out-of-range second tile."""
    import math
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(B * H + cin)
    x = torch.randn(B, H, W, cin, generator=g).half().cuda()
    outs = []
    for max_ctas in (0, 148):
        if kind == 'convt':
            wt = torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(1)) / math.sqrt(cin * 9)
            raw = torch.zeros(B, 2 * H + 2, 2 * W + 2, cout, device='cuda', dtype=torch.float16)
            demod = (1 + 0.1 * torch.randn(B, cout, generator=torch.Generator().manual_seed(2))).cuda()
            op = ops.convt_s2_merged(x, ops.convt_merged_weight(wt.cuda(), 1.0), raw, demod)
            op.desc.max_ctas = max_ctas
            op()
            outs.append(raw)
            continue
        w = (torch.randn(cout, 9 * cin, generator=torch.Generator().manual_seed(3)) / math.sqrt(9 * cin)).half().cuda()
        bias = (0.1 * torch.randn(cout, generator=torch.Generator().manual_seed(4))).cuda()
        out = torch.zeros(B, H, W, cout, device='cuda', dtype=torch.float16)
        kw = dict(bias=bias, act=True, max_ctas=max_ctas)
        if kind == 'mod':
            kw.update(demod=(1 + 0.1 * torch.randn(B, cout, generator=torch.Generator().manual_seed(5))).cuda(),
                      noise=torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(6)).cuda(),
                      noise_gain=torch.full((1,), 0.3).cuda(), noise_strides=(H * W, W))
        if kind == 'res':
            res = torch.randn(B, H, W, cout, generator=torch.Generator().manual_seed(7)).half().cuda()
            kw.update(res=res, res_mode=1, res_strides=(cout, W * cout, H * W * cout), res_wh=(W, H), res_scale=ops.INV_SQRT2)
        ops.conv_same(x, w, out, 3, **kw)()
        outs.append(out)
    torch.cuda.synchronize()
    assert outs[0].float().abs().max().item() > 0.1
    assert torch.equal(outs[0], outs[1])
