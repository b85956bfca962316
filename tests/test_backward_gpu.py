"""GPU parity of the backward pieces of ConvLayer(cin, cout, 3) (image_restoration_b200/backward.py) against torch's
fp32 autograd on the same fp16-rounded operands.

Component tests are strict (each kernel against the fp32 formula on ITS OWN inputs: only fp16 output rounding and the
fp32 summation order differ).  The whole-layer test goes through torch.autograd on both sides and uses a relative RMS
bound, because a pre-activation within ~1e-7 of zero may take the other leaky-ReLU branch on the two sides."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SQRT2 = math.sqrt(2.0)


@pytest.mark.parametrize('n_pix,C', [(1000, 32), (4 * 32 * 96, 256), (7, 512), (12345, 64), (0, 128), (64, 3072), (300, 24)])
def test_lrelu_bias_bwd(n_pix, C):
    from image_restoration_b200 import ops
    torch.manual_seed(C + n_pix)
    dy = torch.randn(n_pix, C, device='cuda').half()
    y = torch.randn(n_pix, C, device='cuda').half()
    y[::3] = 0                                           # y == 0 takes the slope branch (out > 0 is strict in the reference)
    dz, db = ops.lrelu_bias_bwd(dy, y)
    ref = dy.float() * torch.where(y.float() > 0, SQRT2, 0.2 * SQRT2)
    torch.cuda.synchronize()
    assert torch.equal(dz, ref.half())
    ref_b = ref.double().sum(0)
    assert (db.double() - ref_b).abs().max().item() <= 1e-4 * max(1.0, math.sqrt(n_pix))


def test_lrelu_bwd_without_bias_and_bad_channels():
    from image_restoration_b200 import ops
    dy = torch.randn(64, 64, device='cuda').half()
    y = torch.randn(64, 64, device='cuda').half()
    dz, db = ops.lrelu_bias_bwd(dy, y, want_bias=False)
    assert db is None
    assert torch.equal(dz, (dy.float() * torch.where(y.float() > 0, SQRT2, 0.2 * SQRT2)).half())
    with pytest.raises(RuntimeError):
        ops.lrelu_bias_bwd(torch.zeros(4, 12, device='cuda').half(), torch.zeros(4, 12, device='cuda').half())


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 8, 32, 64, 128), (3, 16, 48, 128, 256), (1, 7, 45, 64, 32),
                                              (2, 32, 96, 256, 256), (2, 128, 384, 32, 32)])
def test_dgrad_matches_torch(B, H, W, cin, cout):
    from image_restoration_b200 import ops
    torch.manual_seed(B + H + cin)
    w = (torch.randn(cout, cin, 3, 3, device='cuda') / math.sqrt(9 * cin)).half()
    dz = torch.randn(B, cout, H, W, device='cuda').half()
    ref = torch.nn.grad.conv2d_input((B, cin, H, W), w.float(), dz.float(), padding=1)
    wp = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    dx = torch.empty(B, H, W, cin, device='cuda', dtype=torch.float16)
    ops.conv_dgrad(dz.permute(0, 2, 3, 1).contiguous(), ops.conv_dgrad_weight(wp, cin), dx)()
    torch.cuda.synchronize()
    err = (dx.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f'dgrad B{B} {H}x{W} {cin}->{cout}: max err {err:.3e} of {scale:.3e}')
    assert err <= 1e-3 * scale, (err, scale)             # fp16 rounding of the output: 2^-11 relative


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 16, 48, 64, 128), (2, 32, 96, 256, 256)])
def test_conv_layer_autograd(B, H, W, cin, cout):
    from image_restoration_b200.backward import conv_layer3x3
    torch.manual_seed(cin)
    weight = torch.randn(cout, cin, 3, 3, device='cuda', requires_grad=True)
    bias = (0.1 * torch.randn(cout, device='cuda')).requires_grad_()
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()

    xg = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    y = conv_layer3x3(xg, weight, bias)
    y.backward(dy.permute(0, 2, 3, 1).contiguous())
    got = (y.detach().float().permute(0, 3, 1, 2), xg.grad.float().permute(0, 3, 1, 2), weight.grad.clone(), bias.grad.clone())

    # the reference ConvLayer in fp32 (stylegan2_ocr_arch.py:639-648 + fused_act.py:81-95) on the fp16-rounded operands
    scale = 1.0 / math.sqrt(cin * 9)
    w_ref = ((weight.detach() * scale).half().float() / scale).requires_grad_()
    b_ref = bias.detach().clone().requires_grad_()
    x_ref = x.float().requires_grad_()
    y_ref = F.leaky_relu(F.conv2d(x_ref, w_ref * scale, padding=1) + b_ref.view(1, -1, 1, 1), 0.2) * SQRT2
    y_ref.backward(dy.float())
    ref = (y_ref.detach(), x_ref.grad, w_ref.grad, b_ref.grad)
    torch.cuda.synchronize()
    for name, g, r in zip(('y', 'dx', 'dweight', 'dbias'), got, ref):
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'conv layer {cin}->{cout} {name}: rel rms {rel:.3e}')
        assert rel <= 2e-3, (name, rel)                  # fp16 storage of y, dz and dx: 2^-11 per element


def _fir(x, pad, stride=1):
    """upfirdn2d(x, outer([1,3,3,1]) / 64, pad=(pad, pad)) (+ sampling), the reference's upfirdn2d_native (upfirdn2d.py:162-192)."""
    k = torch.tensor([1., 3., 3., 1.], device=x.device, dtype=x.dtype)
    k2 = (torch.outer(k, k) / 64).view(1, 1, 4, 4).repeat(x.shape[1], 1, 1, 1)
    return F.conv2d(F.pad(x, (pad,) * 4), k2, groups=x.shape[1], stride=stride)


@pytest.mark.parametrize('B,H,W,C', [(2, 8, 24, 64), (1, 32, 96, 256), (3, 6, 10, 32)])
def test_fir_adjoints(B, H, W, C):
    from image_restoration_b200 import ops
    torch.manual_seed(H + C)
    # adjoint of fir_pad22: gradient of sum(fir(x, 2) * d) w.r.t. x
    d = torch.randn(B, C, H + 1, W + 1, device='cuda').half()
    x = torch.zeros(B, C, H, W, device='cuda', requires_grad=True)
    (_fir(x, 2) * d.float()).sum().backward()
    raw = torch.randn(B, H + 2, W + 2, C, device='cuda').half()          # garbage in the padding row / column
    raw[:, :H + 1, :W + 1] = d.permute(0, 2, 3, 1)
    out = torch.empty(B, H, W, C, device='cuda', dtype=torch.float16)
    ops.fir_pad11(raw, out)
    torch.cuda.synchronize()
    err = (out.float().permute(0, 3, 1, 2) - x.grad).abs().max().item()
    assert err <= 2e-3 * x.grad.abs().max().item(), err
    # adjoint of fir_down2 (+ addend)
    d2 = torch.randn(B, C, H // 2, W // 2, device='cuda').half()
    x2 = torch.zeros(B, C, H, W, device='cuda', requires_grad=True)
    (_fir(x2, 1, stride=2) * d2.float()).sum().backward()
    addend = torch.randn(B, H, W, C, device='cuda').half()
    out2 = addend.clone()
    ops.fir_down2_adjoint(d2.permute(0, 2, 3, 1).contiguous(), out2, add=out2)
    ref2 = x2.grad.permute(0, 2, 3, 1) + addend.float()
    torch.cuda.synchronize()
    err2 = (out2.float() - ref2).abs().max().item()
    assert err2 <= 2e-3 * ref2.abs().max().item(), err2


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 16, 48, 64, 128), (2, 32, 96, 256, 256), (1, 32, 96, 32, 64), (2, 8, 24, 64, 256)])
def test_res_block_autograd(B, H, W, cin, cout):
    """ResBlock (stylegan2_ocr_arch.py:708-734) forward + backward against the fp32 restatement through torch.autograd."""
    from image_restoration_b200.backward import ResBlockFunction, res_block
    torch.manual_seed(cin + cout)
    par = dict(w1=torch.randn(cin, cin, 3, 3), b1=0.1 * torch.randn(cin), w2=torch.randn(cout, cin, 3, 3),
               b2=0.1 * torch.randn(cout), ws=torch.randn(cout, cin, 1, 1))
    par = {k: v.cuda().requires_grad_() for k, v in par.items()}
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dout = torch.randn(B, cout, H // 2, W // 2, device='cuda').half()

    xg = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    ResBlockFunction.debug_saved = saved = {}
    out = res_block(xg, par['w1'], par['b1'], par['w2'], par['b2'], par['ws'])
    ResBlockFunction.debug_saved = None
    out.backward(dout.permute(0, 2, 3, 1).contiguous())
    got = {'out': out.detach().float().permute(0, 3, 1, 2), 'dx': xg.grad.float().permute(0, 3, 1, 2)}
    got.update({'d' + k: v.grad.clone() for k, v in par.items()})

    ref_par = {k: v.detach().clone().requires_grad_() for k, v in par.items()}
    s1, s2, ss = 1 / math.sqrt(cin * 9), 1 / math.sqrt(cin * 9), 1 / math.sqrt(cin)
    x_ref = x.float().requires_grad_()
    # The fp16 forward moves pre-activations by ~3e-4, so ~2e-4 of them cross zero relative to an fp32 forward; each
    # crossing changes that element's gradient by a factor 5, which would dominate an RMS comparison (~1e-2).  The
    # reference therefore takes the leaky-ReLU branch from the sign of the kernels' own activations.
    def lrelu(z, act):
        return z * torch.where(act.float().permute(0, 3, 1, 2) > 0, SQRT2, 0.2 * SQRT2)
    t1 = lrelu(F.conv2d(x_ref, ref_par['w1'] * s1, padding=1) + ref_par['b1'].view(1, -1, 1, 1), saved['t1'])
    y2 = lrelu(F.conv2d(_fir(t1, 2), ref_par['w2'] * s2, stride=2) + ref_par['b2'].view(1, -1, 1, 1), saved['y2'])
    sk = F.conv2d(_fir(x_ref, 1), ref_par['ws'] * ss, stride=2)
    out_ref = (y2 + sk) / SQRT2
    out_ref.backward(dout.float())
    ref = {'out': out_ref.detach(), 'dx': x_ref.grad}
    ref.update({'d' + k: v.grad for k, v in ref_par.items()})
    torch.cuda.synchronize()
    for name in ref:
        g, r = got[name], ref[name]
        assert g.shape == r.shape, (name, g.shape, r.shape)
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'res block {cin}->{cout} {H}x{W} {name}: rel rms {rel:.3e}')
        assert rel <= 1.5e-3, (name, rel)    # fp16 weights and fp16 storage of t1, p, y2, dz2, raw, dt1, dx


@pytest.mark.parametrize('B,h,w,C', [(2, 4, 12, 64), (1, 16, 48, 256), (3, 5, 7, 32)])
def test_bilinear_up2_adjoint(B, h, w, C):
    from image_restoration_b200 import ops
    torch.manual_seed(h + C)
    d = torch.randn(B, C, 2 * h, 2 * w, device='cuda').half()
    x = torch.zeros(B, C, h, w, device='cuda', requires_grad=True)
    (F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) * d.float()).sum().backward()
    out = torch.empty(B, h, w, C, device='cuda', dtype=torch.float16)
    ops.bilinear_up2_adjoint(d.permute(0, 2, 3, 1).contiguous(), out, scale=0.5)
    torch.cuda.synchronize()
    err = (out.float().permute(0, 3, 1, 2) - 0.5 * x.grad).abs().max().item()
    assert err <= 2e-3 * x.grad.abs().max().item(), err


@pytest.mark.parametrize('B,h,w,cin,cout', [(2, 8, 24, 256, 256), (2, 16, 48, 256, 64), (1, 32, 96, 64, 32), (2, 4, 12, 128, 128)])
def test_res_up_block_autograd(B, h, w, cin, cout):
    """ResUpBlock (gfpganv1_ocr_arch.py:205-225) forward + backward against the fp32 restatement through torch.autograd
    (leaky-ReLU branches taken from the kernels' own activations, see test_res_block_autograd)."""
    from image_restoration_b200.backward import ResUpBlockFunction, res_up_block
    torch.manual_seed(cin + cout + 1)
    par = dict(w1=torch.randn(cin, cin, 3, 3), b1=0.1 * torch.randn(cin), w2=torch.randn(cout, cin, 3, 3),
               b2=0.1 * torch.randn(cout), ws=torch.randn(cout, cin, 1, 1))
    par = {k: v.cuda().requires_grad_() for k, v in par.items()}
    x = torch.randn(B, cin, h, w, device='cuda').half()
    dout = torch.randn(B, cout, 2 * h, 2 * w, device='cuda').half()

    xg = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    ResUpBlockFunction.debug_saved = saved = {}
    out = res_up_block(xg, par['w1'], par['b1'], par['w2'], par['b2'], par['ws'])
    ResUpBlockFunction.debug_saved = None
    out.backward(dout.permute(0, 2, 3, 1).contiguous())
    got = {'out': out.detach().float().permute(0, 3, 1, 2), 'dx': xg.grad.float().permute(0, 3, 1, 2)}
    got.update({'d' + k: v.grad.clone() for k, v in par.items()})

    ref_par = {k: v.detach().clone().requires_grad_() for k, v in par.items()}
    s3, s1x1 = 1 / math.sqrt(cin * 9), 1 / math.sqrt(cin)
    x_ref = x.float().requires_grad_()

    def lrelu(z, act):
        return z * torch.where(act.float().permute(0, 3, 1, 2) > 0, SQRT2, 0.2 * SQRT2)

    def up(t):
        return F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False)
    t1 = lrelu(F.conv2d(x_ref, ref_par['w1'] * s3, padding=1) + ref_par['b1'].view(1, -1, 1, 1), saved['t1'])
    y2 = lrelu(F.conv2d(up(t1), ref_par['w2'] * s3, padding=1) + ref_par['b2'].view(1, -1, 1, 1), saved['y2'])
    sk = F.conv2d(up(x_ref), ref_par['ws'] * s1x1)
    out_ref = (y2 + sk) / SQRT2
    out_ref.backward(dout.float())
    ref = {'out': out_ref.detach(), 'dx': x_ref.grad}
    ref.update({'d' + k: v.grad for k, v in ref_par.items()})
    torch.cuda.synchronize()
    for name in ref:
        g, r = got[name], ref[name]
        assert g.shape == r.shape, (name, g.shape, r.shape)
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'res up block {cin}->{cout} {h}x{w} {name}: rel rms {rel:.3e}')
        assert rel <= 1.5e-3, (name, rel)


def test_conv_without_activation_autograd():
    """Second conv of an SFT head (EqualConv2d 3x3 with bias, no activation; gfpganv1_ocr_arch.py:322-339)."""
    from image_restoration_b200.backward import conv_layer3x3
    torch.manual_seed(5)
    B, H, W, cin, cout = 2, 16, 48, 128, 64
    weight = torch.randn(cout, cin, 3, 3, device='cuda', requires_grad=True)
    bias = torch.ones(cout, device='cuda', requires_grad=True)
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()
    xg = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    y = conv_layer3x3(xg, weight, bias, False)
    y.backward(dy.permute(0, 2, 3, 1).contiguous())
    scale = 1.0 / math.sqrt(cin * 9)
    w_ref, b_ref, x_ref = weight.detach().clone().requires_grad_(), bias.detach().clone().requires_grad_(), x.float().requires_grad_()
    y_ref = F.conv2d(x_ref, w_ref * scale, b_ref, padding=1)
    y_ref.backward(dy.float())
    torch.cuda.synchronize()
    for name, g, r in (('y', y.detach().float().permute(0, 3, 1, 2), y_ref.detach()), ('dx', xg.grad.float().permute(0, 3, 1, 2), x_ref.grad),
                       ('dweight', weight.grad, w_ref.grad), ('dbias', bias.grad, b_ref.grad)):
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'conv (no activation) {name}: rel rms {rel:.3e}')
        assert rel <= 1e-3, (name, rel)


@pytest.mark.parametrize('B,cin,cout', [(64, 12288, 3072), (8, 256, 512), (3, 64, 128)])
def test_equal_linear_autograd(B, cin, cout):
    """final_linear of GFPGANv1OCR (EqualLinear, stylegan2_ocr_arch.py:165-175) forward + backward."""
    from image_restoration_b200.backward import equal_linear
    torch.manual_seed(B)
    weight = torch.randn(cout, cin, device='cuda', requires_grad=True)
    bias = (0.1 * torch.randn(cout, device='cuda')).requires_grad_()
    x = torch.randn(B, cin, device='cuda').half()
    dy = torch.randn(B, cout, device='cuda').half()
    xg = x.clone().requires_grad_()
    y = equal_linear(xg, weight, bias)
    y.backward(dy)
    scale = 1.0 / math.sqrt(cin)
    w_ref, b_ref, x_ref = weight.detach().clone().requires_grad_(), bias.detach().clone().requires_grad_(), x.float().requires_grad_()
    y_ref = F.linear(x_ref, w_ref * scale, b_ref)
    y_ref.backward(dy.float())
    torch.cuda.synchronize()
    for name, g, r in (('y', y.detach().float(), y_ref.detach()), ('dx', xg.grad.float(), x_ref.grad), ('dweight', weight.grad, w_ref.grad),
                       ('dbias', bias.grad, b_ref.grad)):
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'equal linear {cin}->{cout} B{B} {name}: rel rms {rel:.3e}')
        assert rel <= 1e-3, (name, rel)


def test_unet_forward_backward_against_oracle():
    """Trainable part of GFPGANv1OCR (U-Net encoder -> style code, decoder -> SFT conditions; everything optimizer_g updates
    with fix_decoder=True) through backward.unet_forward, against torch.autograd over the fp32 oracle on the same seeded
    stock-init parameters.  The fp16 forward flips the leaky-ReLU branch of ~2e-4 of the pre-activations relative to an
    fp32 forward (each flip changes that element's gradient five-fold), so the end-to-end gradient bound is a few per cent
    relative RMS plus a cosine; the per-block tests above pin every kernel to 5e-4 with the branches held equal."""
    from image_restoration_b200 import GFPGANv1OCR
    from image_restoration_b200.backward import unet_forward
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    from tests.helpers import KW
    torch.manual_seed(0)
    W, H, B = 384, 128, 2
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW)
    sd = {k: v.detach().clone().cuda() for k, v in net.state_dict().items()}
    trainable = [k for k in sd if k.split('.')[0] in ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up',
                                                       'condition_scale', 'condition_shift')]
    sd_a = {k: (v.clone().requires_grad_() if k in trainable else v) for k, v in sd.items()}
    sd_b = {k: (v.clone().requires_grad_() if k in trainable else v) for k, v in sd.items()}
    x = (torch.rand(B, 3, H, W, device='cuda') * 2 - 1)

    style, conds = unet_forward(sd_a, x, different_w=True, num_style_feat=KW['num_style_feat'])
    cfg = OcrNetConfig(input_width=W, input_height=H, **KW)
    taps = {}
    gfpgan_ocr_forward.__wrapped__(sd_b, cfg, x, return_rgb=False, taps=taps)    # the oracle without its torch.no_grad()
    L = cfg.num_levels
    ref_conds = [taps[f'{n}{i}'] for i in range(L) for n in ('scale', 'shift')]
    g = torch.Generator(device='cuda').manual_seed(1)
    loss_a = loss_b = 0.0
    cot = torch.randn(style.shape, device='cuda', generator=g).half()
    loss_a = loss_a + (style.float() * cot.float()).sum()
    loss_b = loss_b + (taps['style_code'] * cot.float()).sum()
    fwd = [('style_code', style.detach().float(), taps['style_code'].detach())]
    for i, (c, r) in enumerate(zip(conds, ref_conds)):
        cot = torch.randn(c.shape, device='cuda', generator=g).half()
        loss_a = loss_a + (c.float() * cot.float()).sum()
        loss_b = loss_b + (r * cot.float().permute(0, 3, 1, 2)).sum()
        fwd.append((f'cond{i}', c.detach().float().permute(0, 3, 1, 2), r.detach()))
    loss_a.backward()
    loss_b.backward()
    torch.cuda.synchronize()
    for name, a, r in fwd:
        rel = ((a - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'unet forward {name}: rel rms {rel:.3e}')
        assert rel <= 5e-3, (name, rel)
    worst = 0.0
    for k in trainable:
        ga, gb = sd_a[k].grad, sd_b[k].grad
        assert ga is not None and gb is not None and ga.shape == gb.shape, k
        rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt().clamp_min(1e-30)).item()
        cos = F.cosine_similarity(ga.double().flatten(), gb.double().flatten(), dim=0).item()
        worst = max(worst, rel)
        print(f'unet grad {k}: rel rms {rel:.3e} cos {cos:.6f}')
        assert rel <= 4e-2 and cos >= 0.999, (k, rel, cos)
    print(f'unet backward: {len(trainable)} parameter tensors, worst rel rms {worst:.3e}')
