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


@pytest.mark.parametrize('n_pix,C', [(1000, 32), (4 * 32 * 96, 256), (7, 512), (12345, 64), (0, 128)])
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
        ops.lrelu_bias_bwd(torch.zeros(4, 24, device='cuda').half(), torch.zeros(4, 24, device='cuda').half())


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
