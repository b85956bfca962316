"""GPU parity of the conv weight-gradient kernel (b200ir_conv_wgrad, tcgen05 GEMM over pixels with MN-major operands)
against torch's fp32 conv2d weight gradient computed from the same fp16-rounded operands.  Products of fp16 values are
exact in fp32 and the accumulation is fp32, so only the summation order differs: rel 1e-3 of the largest entry
(observed ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 8, 32, 64, 128), (3, 16, 48, 128, 256), (1, 7, 45, 64, 128),
                                              (4, 32, 96, 256, 256), (2, 4, 12, 512, 512), (2, 16, 48, 64, 64), (1, 32, 96, 32, 32),
                                              (2, 8, 24, 32, 64), (1, 8, 24, 192, 24), (1, 8, 16, 48, 200)])
def test_wgrad_matches_torch(B, H, W, cin, cout):
    from image_restoration_b200 import ops
    torch.manual_seed(B * 100 + H)
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, 3, 3), dy.float(), padding=1)     # [cout, cin, 3, 3]
    ref = ref.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    got = ops.conv_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f'wgrad B{B} {H}x{W} {cin}->{cout}: max err {err:.3e} of {scale:.3e}')
    assert err <= 1e-3 * scale, (err, scale)


def test_wgrad_rejects_unsupported_channels():
    from image_restoration_b200 import ops
    x = torch.zeros(1, 8, 8, 24, device='cuda', dtype=torch.float16)          # cin must be a multiple of 16
    dy = torch.zeros(1, 8, 8, 128, device='cuda', dtype=torch.float16)
    with pytest.raises(RuntimeError):
        ops.conv_wgrad(x, dy)


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 8, 32, 64, 128), (3, 16, 48, 256, 256), (1, 5, 37, 128, 128)])
def test_wgrad_1x1(B, H, W, cin, cout):
    from image_restoration_b200 import ops
    torch.manual_seed(B + W)
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, 1, 1), dy.float())[:, :, 0, 0]
    got = ops.conv1x1_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    err, scale = (got - ref).abs().max().item(), ref.abs().max().item()
    print(f'wgrad 1x1 B{B} {H}x{W} {cin}->{cout}: max err {err:.3e} of {scale:.3e}')
    assert err <= 1e-3 * scale, (err, scale)


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 8, 32, 64, 128), (2, 32, 96, 64, 256), (3, 16, 48, 256, 256), (1, 6, 70, 128, 128),
                                              (2, 16, 48, 32, 64)])
def test_wgrad_3x3_stride2(B, H, W, cin, cout):
    """ResBlock.conv2: F.conv2d(p, W, stride=2) over the (H+1)x(W+1) FIR output held in a [B,H+2,W+2,C] buffer."""
    from image_restoration_b200 import ops
    torch.manual_seed(B + W + 1)
    p = torch.randn(B, H + 2, W + 2, cin, device='cuda').half()          # padding row / column holds garbage on purpose
    dy = torch.randn(B, cout, H // 2, W // 2, device='cuda').half()
    valid = p[:, :H + 1, :W + 1].permute(0, 3, 1, 2).float()
    ref = torch.nn.grad.conv2d_weight(valid, (cout, cin, 3, 3), dy.float(), stride=2)
    ref = ref.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    got = ops.conv3x3_s2_wgrad(p, H, W, dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    err, scale = (got - ref).abs().max().item(), ref.abs().max().item()
    print(f'wgrad 3x3 s2 B{B} {H}x{W} {cin}->{cout}: max err {err:.3e} of {scale:.3e}')
    assert err <= 1e-3 * scale, (err, scale)


@pytest.mark.parametrize('fold', [0, 1])
@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 16, 48, 32, 32), (1, 9, 20, 32, 64), (2, 8, 24, 64, 64), (1, 16, 40, 64, 32),
                                              (1, 8, 18, 32, 32)])
def test_wgrad_pixel_fold_modes(B, H, W, cin, cout, fold, monkeypatch):
    """Low-channel layers on pixel-folded views (ops.wgrad_fold + b200ir_wgrad_unfold): every fold mode gives the same weight
    gradient as torch (mode 0 = no folding; W = 18 is not a multiple of 4, so 32 -> 32 stays unfolded there)."""
    from image_restoration_b200 import ops
    monkeypatch.setattr(ops, '_WGRAD_FOLD', fold)
    torch.manual_seed(W + cin)
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, 3, 3), dy.float(), padding=1).permute(0, 2, 3, 1).reshape(cout, 9, cin)
    got = ops.conv_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    err, scale = (got - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 1e-3 * scale, (fold, ops.wgrad_fold(cin, cout, W), err, scale)
